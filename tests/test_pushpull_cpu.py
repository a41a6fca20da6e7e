"""Periodic push-pull anti-entropy ([U] memberlist/state.go pushPull, serf/delegate.go
LocalState / MergeRemoteState), opt-in with GSIM_FLAG_PUSH_PULL: the kernel's row body (hostemu)
against the oracle, and the property that makes it "the other half of convergence": a broadcast
that gossip alone leaves stranded under loss is completed by push-pull."""
import pytest

import scenarios as sc
from consul_b200.pool import (FLAG_PUSH_PULL, NEVER, PRED_RUMOR_CONVERGED, Pool, consul_test_config,
                              lan_config, wan_config)
from oracle_binding import OraclePool
from parity import compare_pools

SEC = 1_000_000_000


@pytest.fixture()
def make(hostemu_lib):
    return lambda cfg: [Pool(cfg, hostemu_lib), OraclePool(cfg)]


def stranded_event(make, lib, flags, n=3000, ticks=900, every=25):
    """50 % loss and a retransmit budget of 4: gossip reaches only part of the cluster."""
    cfg = lan_config(lib, capacity=n + 2, n_initial=n, seed=0x5EED00AA, packet_loss_ppm=500000,
                     retransmit_mult=1, flags=flags, push_pull_interval_ns=1 * SEC)
    pools = make(cfg)
    slot = sc.both(pools, lambda p: p.user_event(17, b"deploy", b"v1", False))
    x = sc.both(pools, lambda p: p.member_add())
    sc.both(pools, lambda p: p.join(x, [3]))
    for p in pools:
        p.crash_many([100, 200])                     # push-pull partners that never answer
    sc.step_compare(pools, ticks, every, f"stranded flags={flags}")
    return pools, slot, x


def test_push_pull_completes_what_gossip_strands(make, hostemu_lib):
    pools, slot, x = stranded_event(make, hostemu_lib, 0)
    info = pools[0].rumor_info(slot)
    assert info["queued_count"] == 0 and info["heard_count"] < pools[0].stats()["n_up"]   # stranded
    stranded = info["heard_count"]
    assert pools[0].stats()["push_pulls"] == 0

    pools, slot, x = stranded_event(make, hostemu_lib, FLAG_PUSH_PULL)
    for p in pools:
        s = p.stats()
        info = p.rumor_info(slot)
        assert info["heard_count"] == s["n_up"] > stranded          # everyone alive got it
        assert info["converged_tick"] != NEVER
        assert s["push_pulls"] > s["n_up"]                           # several rounds (80-tick period)
        assert p.num_nodes(x) == s["n_members"]                      # the joiner learnt everybody
        assert int(p.column("ltime_event")[: s["n_members"]][p.column("key")[: s["n_members"]] & 3 == 1].min()) >= 2


def test_push_pull_period_and_stat(make, hostemu_lib):
    """pushPullScale: 30 s up to 32 members, x8 at 3000; one exchange per member per period."""
    n = 3000
    cfg = lan_config(hostemu_lib, capacity=n, n_initial=n, seed=5, flags=FLAG_PUSH_PULL, push_pull_interval_ns=2 * SEC)
    pools = make(cfg)
    period = 8 * 20                                                   # ceil(log2(3000) - 5) + 1 = 8
    sc.step_compare(pools, 2 * period, period // 2, "two periods")
    for p in pools:
        assert p.stats()["push_pulls"] == 2 * n
        assert p.stats()["rumors_accepted"] == 0 and p.stats()["suspects"] == 0


def test_push_pull_small_cluster_per_member_phases(make, hostemu_lib):
    """3 agents, Consul's test timing, per-member ticker phases (phase_group = 1)."""
    cfg = consul_test_config(hostemu_lib, capacity=8, n_initial=0, seed=2, flags=1 | FLAG_PUSH_PULL,
                             phase_group=1, push_pull_interval_ns=500_000_000, packet_loss_ppm=300000)
    pools = make(cfg)
    ids = [sc.both(pools, lambda p: p.member_add(watched=True)) for _ in range(3)]
    sc.both(pools, lambda p: p.join(1, [0]))
    sc.both(pools, lambda p: p.join(2, [0]))
    sc.both(pools, lambda p: p.user_event(2, b"e", b"", False))
    sc.step_compare(pools, 200, 1, "small")
    for p in pools:
        assert p.stats()["push_pulls"] > 10
        for obs in ids:
            assert len(p.members(obs)) == 3


def test_push_pull_with_wan_latency(make, hostemu_lib):
    from consul_b200.wan import c5_latency_matrix
    n = 4096
    cfg = wan_config(hostemu_lib, capacity=n, n_initial=n, seed=8, flags=FLAG_PUSH_PULL, mailbox_depth=8,
                     push_pull_interval_ns=2 * SEC, packet_loss_ppm=400000, retransmit_mult=1)
    pools = make(cfg)
    for p in pools:
        p.latency_set(c5_latency_matrix(32))
    slot = sc.both(pools, lambda p: p.user_event(0, b"x", b"y", False))
    sc.step_compare(pools, 400, 20, "wan + push-pull")
    t = sc.both(pools, lambda p: p.run_until(PRED_RUMOR_CONVERGED, slot, 3000, 50))
    assert t != NEVER
    compare_pools(*pools, "converged")


def test_snapshot_restore_mid_exchange(hostemu_lib):
    n = 2000
    cfg = lan_config(hostemu_lib, capacity=n, n_initial=n, seed=4, flags=FLAG_PUSH_PULL,
                     push_pull_interval_ns=1 * SEC, packet_loss_ppm=500000, retransmit_mult=1)
    p = Pool(cfg, hostemu_lib)
    slot = p.user_event(1, b"a", b"b", False)
    p.step(101)                                                       # requests/answers in flight
    blob = p.snapshot()
    p.step(300)
    q = Pool(cfg, hostemu_lib)
    q.restore(blob)
    q.step(300)
    assert q.state_hash() == p.state_hash() and q.rumor_info(slot) == p.rumor_info(slot)
    with pytest.raises(Exception):
        Pool(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=4), hostemu_lib).restore(blob)


@pytest.mark.parametrize("order", ["1", "2"])
def test_row_order_independence(order):
    """Requests, answers and clock merges are commutative mailboxes: reverse / odd-even row order
    gives the oracle's result too."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys; sys.path[:0]=[%r,%r]\n"
        "import test_pushpull_cpu as t\n"
        "from consul_b200 import _lib\n"
        "from consul_b200.pool import Pool, FLAG_PUSH_PULL\n"
        "from oracle_binding import OraclePool\n"
        "L=_lib.load(%r)\n"
        "mk=lambda cfg:[Pool(cfg,L),OraclePool(cfg)]\n"
        "t.stranded_event(mk,L,FLAG_PUSH_PULL,n=1500,ticks=500)\n"
        "t.test_push_pull_small_cluster_per_member_phases(mk,L)\n"
    ) % (os.path.dirname(here), here, os.path.join(here, "hostemu", "libgsim_hostemu.so"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GSIM_HOSTEMU_ORDER=order),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
