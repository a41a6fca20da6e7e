"""Host-side checks that need no GPU: the C ABI library loads and exports every symbol the
header declares, refuses to run without a device, and the Python mirror matches the header."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "gsim.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsim_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from consul_b200 import _lib
    lib = C.CDLL(_lib.DEFAULT_LIB)
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"libgsim.so does not export {n}"
    bound = {s[0] for s in _lib.SIGNATURES}
    assert set(names) == bound, set(names) ^ bound


def test_config_struct_matches_header():
    from consul_b200 import _lib
    from consul_b200.pool import lan_config
    cfg = lan_config(_lib.lib())
    assert cfg.struct_size == C.sizeof(_lib.GsimConfig)      # the library wrote its own sizeof
    assert _lib.lib().gsim_abi_version() == 1


def test_no_cpu_fallback():
    """Without a B200 the product must fail loudly (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from consul_b200.pool import GsimError, Pool, lan_config
    with pytest.raises(GsimError) as e:
        Pool(lan_config(capacity=16, n_initial=4))
    assert e.value.code == -2


def test_product_never_touches_the_oracle():
    """consul_b200/ must not import, link or load anything under oracle/ or tests/."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "consul_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"liboracle|oracle_binding|oracle/|hostemu_backend|libgsim_hostemu", txt):
                    # comments that merely mention the test-only build are fine in gs_backend.h
                    if f == "gs_backend.h" or f == "_lib.py":
                        continue
                    bad.append(f)
    assert not bad, bad


def test_invalid_arguments(hostemu_lib):
    from consul_b200.pool import GsimError, Pool, lan_config
    for kw in (dict(capacity=0), dict(capacity=4, n_initial=5), dict(phase_group=100),
               dict(probe_timeout_ns=10**9), dict(gossip_nodes=9)):
        with pytest.raises(GsimError):
            Pool(lan_config(hostemu_lib, **{**dict(capacity=8, n_initial=2), **kw}), hostemu_lib)
    p = Pool(lan_config(hostemu_lib, capacity=4, n_initial=4), hostemu_lib)
    with pytest.raises(GsimError) as e:
        p.member_add()                                            # capacity exhausted
    assert e.value.code == -4
    with pytest.raises(GsimError) as e:
        p.user_event(0, b"n" * 300, b"p" * 300)                    # > UserEventSizeLimit
    assert e.value.code == -7
    with pytest.raises(GsimError):
        p.join(9, [0])
    p.crash(1)
    with pytest.raises(GsimError):
        p.user_event(1, b"x", b"y")                                # crashed members cannot fire
    p.step(0)
    assert p.now == 0


def test_snapshot_roundtrip_on_host(hostemu_lib):
    from consul_b200.pool import Pool, lan_config
    p = Pool(lan_config(hostemu_lib, capacity=3001, n_initial=3000, seed=5, packet_loss_ppm=80000), hostemu_lib)
    x = p.member_add()
    p.join(x, [1])
    p.user_event(2, b"ev", b"payload")
    p.crash_fraction(30000, 1)
    p.step(40)
    blob = p.snapshot()
    p.step(100)
    want = (p.state_hash(), p.stats())
    p.step(33)
    p.restore(blob)
    assert p.now == 40
    p.step(100)
    assert (p.state_hash(), p.stats()) == want
    from parity import active_mask
    slots = [r for r in range(30) if (active_mask(p) >> r) & 1 and p.rumor_info(r)["kind"] == 4]
    assert [p.user_event_get(r) for r in slots] == [(b"ev", b"payload")]   # payload survives restore
