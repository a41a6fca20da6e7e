"""CSR peer graph (BASELINE north_star: "message-passing kernel over a CSR peer graph"):
  * a CSR whose every row is 0..n-1 reproduces the complete-graph results bit for bit;
  * on a restricted topology (ring of segments) the kernel body equals the oracle, a broadcast
    never leaves the reachable component and takes longer than on the complete graph;
  * the probe ring of a member visits exactly its row."""
import numpy as np
import pytest

import scenarios as sc
from consul_b200.pool import NEVER, PRED_CRASHED_ALL_DEAD, PRED_RUMOR_CONVERGED, Pool, lan_config
from oracle_binding import OraclePool
from parity import compare_columns, compare_pools, compare_stats


@pytest.fixture()
def make(hostemu_lib):
    return lambda cfg: [Pool(cfg, hostemu_lib), OraclePool(cfg)]


def complete_csr(n):
    return np.arange(0, n * n + 1, n, dtype=np.uint32), np.tile(np.arange(n, dtype=np.uint32), n)


def segment_ring_csr(n, seg):
    """Members in segments of `seg`; a member knows its own segment and the next one (a ring)."""
    rows = []
    nseg = (n + seg - 1) // seg
    for i in range(n):
        s = i // seg
        own = np.arange(s * seg, min(n, (s + 1) * seg))
        nxt = ((s + 1) % nseg)
        other = np.arange(nxt * seg, min(n, (nxt + 1) * seg))
        rows.append(np.unique(np.concatenate([own, other])))
    rp = np.zeros(n + 1, dtype=np.uint32)
    rp[1:] = np.cumsum([len(r) for r in rows])
    return rp, np.concatenate(rows).astype(np.uint32)


def script(p):
    slot = p.user_event(3, b"deploy", b"v1", False)
    p.crash_many([40, 41])
    p.step(150)
    return slot


@pytest.mark.parametrize("which", ["kernel-body", "oracle"])
def test_complete_csr_equals_implicit_complete_graph(which, hostemu_lib):
    n = 600
    cfg = lan_config(hostemu_lib, capacity=n, n_initial=n, seed=9, packet_loss_ppm=100000)
    mk = (lambda: Pool(cfg, hostemu_lib)) if which == "kernel-body" else (lambda: OraclePool(cfg))
    a, b = mk(), mk()
    b.graph_set(*complete_csr(n))
    script(a)
    script(b)
    compare_stats(a, b, "implicit vs CSR")
    compare_columns(a, b, "implicit vs CSR")
    assert a.state_hash() == b.state_hash()
    assert len(b.members(5)) == n


def test_segment_ring_parity_and_reachability(make, hostemu_lib):
    n, seg = 1024, 128
    cfg = lan_config(hostemu_lib, capacity=n, n_initial=n, seed=13)
    pools = make(cfg)
    rp, ci = segment_ring_csr(n, seg)
    for p in pools:
        p.graph_set(rp, ci)
        assert len(p.members(0)) == 2 * seg                  # own segment + next one
    slot = sc.both(pools, lambda p: p.user_event(0, b"e", b"p", False))
    sc.step_compare(pools, 20, 1, "early")
    t = sc.both(pools, lambda p: p.run_until(PRED_RUMOR_CONVERGED, slot, 2000, 4))
    assert t != NEVER                                         # the ring of segments is connected
    compare_pools(*pools, "converged on the ring")
    # the same event on the complete graph arrives sooner (8 segments = 7 hops on the ring)
    q = Pool(cfg, hostemu_lib)
    s2 = q.user_event(0, b"e", b"p", False)
    assert q.run_until(PRED_RUMOR_CONVERGED, s2, 2000, 4) < t
    # failure detection works along the edges too
    for p in pools:
        p.crash_many([300, 700])
    td = sc.both(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 4000, 25))
    assert td != NEVER
    compare_pools(*pools, "crashed dead on the ring")
    for p in pools:
        assert p.stats()["refutes"] == 0 and p.stats()["deads"] == 2


def test_disconnected_components_do_not_leak(make, hostemu_lib):
    """Two halves with no edge between them: a broadcast stays in its half (and never converges)."""
    n = 512
    half = n // 2
    rows = [np.arange(0, half) if i < half else np.arange(half, n) for i in range(n)]
    rp = np.zeros(n + 1, dtype=np.uint32)
    rp[1:] = np.cumsum([len(r) for r in rows])
    ci = np.concatenate(rows).astype(np.uint32)
    cfg = lan_config(hostemu_lib, capacity=n, n_initial=n, seed=5)
    pools = make(cfg)
    for p in pools:
        p.graph_set(rp, ci)
    slot = sc.both(pools, lambda p: p.user_event(7, b"x", b"", False))
    sc.step_compare(pools, 200, 20, "two islands")
    for p in pools:
        heard = (p.column("heard")[:n] >> slot) & 1
        assert heard[:half].all() and not heard[half:].any()
        assert p.rumor_info(slot)["converged_tick"] == NEVER
        assert p.stats()["suspects"] == 0


def test_probe_ring_visits_exactly_the_row(hostemu_lib):
    """One pass of the probe ring of member 0 targets each member of its row once."""
    n = 256
    row0 = np.array([0, 5, 9, 17, 33, 65, 129, 200], dtype=np.uint32)          # includes itself
    rows = [row0] + [np.array([0, i], dtype=np.uint32) for i in range(1, n)]
    rp = np.zeros(n + 1, dtype=np.uint32)
    rp[1:] = np.cumsum([len(r) for r in rows])
    ci = np.concatenate(rows).astype(np.uint32)
    # everybody but member 0's row is crashed => every probe of member 0 is a (failed) probe whose
    # target we can read back from probe_tgt
    cfg = lan_config(hostemu_lib, capacity=n, n_initial=n, seed=3, disable_tcp_pings=1, indirect_checks=0)
    for mk in (lambda: Pool(cfg, hostemu_lib), lambda: OraclePool(cfg)):
        p = mk()
        p.graph_set(rp, ci)
        p.crash_many([int(x) for x in row0[1:]])
        seen = []
        for _ in range(400):
            p.step(1)
            meta0 = int(p.column("meta")[0])
            if (meta0 >> 3) & 3 == 1:                                          # WAIT_T: a probe is in flight
                tgt = int(p.column("probe_tgt")[0])
                if not seen or seen[-1] != tgt:
                    seen.append(tgt)
            if len(seen) >= 7:
                break
        assert sorted(seen[:7]) == sorted(int(x) for x in row0[1:]), seen


def test_graph_api_contract(hostemu_lib):
    n = 64
    cfg = lan_config(hostemu_lib, capacity=n + 4, n_initial=n, seed=1)
    p = Pool(cfg, hostemu_lib)
    rp, ci = complete_csr(n)
    with pytest.raises(Exception):
        p.graph_set(rp[:-1], ci)                               # one row short
    bad = ci.copy()
    bad[3] = n + 7
    with pytest.raises(Exception):
        p.graph_set(rp, bad)                                   # edge to a member that does not exist
    p.graph_set(rp, ci)
    with pytest.raises(Exception):
        p.member_add()                                         # static topology
    p.graph_set(None, None)
    assert p.member_add() == n
