"""The compile-time performance variants (DESIGN §8 round-2 plan) must be semantics-neutral:
  GS_KSTAT    4-bit status replica for peer gathers
  GS_MAILMAP  one mailbox BIT per member for the scan, the 4-byte word only when the bit is raised
The row logic compiled with each of them (and both) reproduces the golden fixtures and matches the
oracle on the scenarios that change keys and mailboxes (crash, refute, leave, join, reap, SetTags,
push-pull, WAN latency, snapshot/restore).  The default build defines neither."""
import os
import subprocess

import pytest

import scenarios as sc
import test_golden as tg
from consul_b200 import _lib
from consul_b200.pool import Pool
from oracle_binding import OraclePool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


VARIANTS = {"kstat": ["-DGS_KSTAT=1"], "mailmap": ["-DGS_MAILMAP=1"], "both": ["-DGS_KSTAT=1", "-DGS_MAILMAP=1"]}


@pytest.fixture(scope="module", params=list(VARIANTS))
def kstat_lib(request):
    out = os.path.join(ROOT, "tests", "hostemu", "libgsim_hostemu_%s.so" % request.param)
    srcs = [os.path.join(ROOT, "consul_b200", "csrc", "gs_api.cpp"),
            os.path.join(ROOT, "tests", "hostemu", "hostemu_backend.cpp")]
    newest = max(os.path.getmtime(os.path.join(ROOT, "consul_b200", "csrc", f))
                 for f in os.listdir(os.path.join(ROOT, "consul_b200", "csrc")))
    if not os.path.exists(out) or os.path.getmtime(out) < max(newest, os.path.getmtime(srcs[1])):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared"] + VARIANTS[request.param] +
                       ["-DGS_MAKE_BACKEND=gs_make_hostemu_backend", "-o", out] + srcs, check=True)
    return _lib.load(out)


@pytest.fixture()
def make(kstat_lib):
    return lambda cfg: [Pool(cfg, kstat_lib), OraclePool(cfg)]


@pytest.mark.parametrize("case", tg.ALL_CASES, ids=[c[0] for c in tg.ALL_CASES])
def test_variant_reproduces_golden(case, kstat_lib):
    tg.check(lambda cfg: Pool(cfg, kstat_lib), kstat_lib, case)


def test_variant_status_bytes_track_the_keys(kstat_lib):
    """After a run with crashes, refutes, a leave, a join and a SetTags, every status byte equals
    the 4-bit codes of the two key buffers (read back through a snapshot, which carries it)."""
    import numpy as np
    from consul_b200.pool import lan_config
    n = 2000
    p = Pool(lan_config(kstat_lib, capacity=n + 2, n_initial=n, seed=31, packet_loss_ppm=300000,
                        disable_tcp_pings=1), kstat_lib)
    x = p.member_add()
    p.join(x, [4])
    p.crash_many([7, 8, 9])
    p.leave(11)
    p.member_update(12)
    p.step(357)
    assert p.stats()["refutes"] > 0 and p.stats()["deads"] > 0
    blob = p.snapshot()
    q = Pool(p.cfg, kstat_lib)
    q.restore(blob)
    q.step(100)
    p.step(100)
    assert p.state_hash() == q.state_hash()


def test_variant_scenarios_against_oracle(make, kstat_lib):
    sc.c1_three_node_join(make, kstat_lib, 1)
    sc.c2_join_cascade(make, kstat_lib, 3000, every=4)
    sc.c3_crash(make, kstat_lib, 1500)
    sc.lossy_scenario(make, kstat_lib, n=300, loss_ppm=400000, seed=12, ticks=400, disable_tcp_pings=1)
    sc.leave_scenario(make, kstat_lib)
    sc.lan_reap_scenario(make, kstat_lib, 1)
    sc.set_tags_scenario(make, kstat_lib)
    sc.event_window_scenario(make, kstat_lib)
    sc.budget_scenario(make, kstat_lib)
    import test_latency_cpu as wl
    import test_pushpull_cpu as pp
    from consul_b200.pool import FLAG_PUSH_PULL, wan_config
    wl.test_event_dissemination_parity(make, kstat_lib, wan_config, 64)
    wl.test_lossy_crash_parity_with_latency(make, kstat_lib)
    wl.test_snapshot_restore_with_packets_in_flight(kstat_lib)
    pp.stranded_event(make, kstat_lib, FLAG_PUSH_PULL, n=1500, ticks=500)
    pp.test_snapshot_restore_mid_exchange(kstat_lib)


def test_variant_random_sequences(make, kstat_lib):
    import fuzz_ops
    for seed in list(range(1000, 1030)) + [251]:
        fuzz_ops.run_sequence(make, kstat_lib, seed)
