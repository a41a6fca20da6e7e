"""What the sm_100a build of libgsim.so contains, read from its SASS here (nvcc cross-compiles without a
GPU, cuobjdump disassembles without one): the Blackwell-side mechanisms DESIGN.md names are in the binary the
GPU box will load — bulk-copy scan with mbarrier completion, system-scope reductions for sharded mailbox
posts, both instantiations of the window kernel — and nothing was built for another architecture."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "consul_b200", "libgsim.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name:
            funcs[name].append(line)
    return {k: "\n".join(v) for k, v in funcs.items()}


def test_only_sm_100a_code_is_in_the_library():
    elfs = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elfs))
    assert archs == {"sm_100a"}, elfs


def test_tick_kernel_scans_with_bulk_copies_and_mbarriers(sass):
    tick = [v for k, v in sass.items() if "gs_tick_kernel" in k]
    assert len(tick) == 2                                      # with and without network coordinates
    for body in tick:
        assert "UBLKCP" in body                                # cp.async.bulk (1-D TMA) global -> shared
        assert "SYNCS.ARRIVE.TRANS64" in body and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in body   # mbarrier expect_tx / try_wait
        assert "LDGSTS" not in body                            # the per-lane cp.async ring of round 1 is gone


def test_sharded_mailbox_posts_are_system_scope_reductions(sass):
    tick = "\n".join(v for k, v in sass.items() if "gs_tick_kernel" in k or "gs_row_step_call" in k)
    assert "REDG.E.OR.STRONG.SYS" in tick                      # red.global.sys.or.b32 (default)
    assert "ATOMG.E.OR.STRONG.SYS" in tick                     # atom.global.sys.or (GSIM_FLAG_SHARD_ATOM)
    assert "ATOMG.E.MIN.64.STRONG.SYS" in tick                 # accusation chains


def test_window_kernel_has_a_closed_form_instantiation_without_the_probe_loop(sass):
    win = {k: v for k, v in sass.items() if "gs_window_kernel" in k}
    assert len(win) == 4                                       # coordinates x pristine
    closed = [v for k, v in win.items() if re.search(r"gs_window_kernelILb[01]ELb1E", k)]
    loop = [v for k, v in win.items() if re.search(r"gs_window_kernelILb[01]ELb0E", k)]
    assert len(closed) == 2 and len(loop) == 2
    # (an entry's listing includes the out-of-line generic path it calls, which gathers too: compare counts)
    gathers = lambda body: body.count("LDG.E.U8.STRONG.GPU")   # the byte-wide gather of a target's status  # noqa: E731
    for c, l in zip(sorted(closed, key=gathers), sorted(loop, key=gathers)):
        assert gathers(c) < gathers(l), (gathers(c), gathers(l))   # the four lock-step gathers of the probe loop are gone
