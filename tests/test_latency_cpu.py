"""WAN latency pools and the two-pool federation of BASELINE config 5, on the
CPU: the kernel's row body (tests/hostemu) against the oracle, bit for bit, plus the properties
the latency model must have (an all-ones matrix is the plain model; latency only ever delays)."""
import numpy as np
import pytest

import scenarios as sc
from consul_b200.pool import (NEVER, PRED_CRASHED_ALL_DEAD, PRED_RUMOR_CONVERGED, Pool, lan_config,
                              wan_config)
from consul_b200.wan import WanFederation, bridge_ids, c5_latency_matrix
from oracle_binding import OraclePool
from parity import compare_pools


def slow_matrix(n_dcs, worst):
    """Asymmetric one-way latencies in [1, worst] whose round trips are not bounded by 5 extra
    ticks (the C5 formula's mods always sum to 0 or 5, so its round trips never exceed WAN's
    ProbeTimeout of 6 ticks)."""
    a = np.arange(n_dcs)[:, None]
    b = np.arange(n_dcs)[None, :]
    m = 1 + (3 * a + 5 * b) % worst
    m[np.arange(n_dcs), np.arange(n_dcs)] = 1
    return m.astype(np.uint8)


@pytest.fixture()
def make(hostemu_lib):
    return lambda cfg: [Pool(cfg, hostemu_lib), OraclePool(cfg)]


def test_c5_matrix_shape():
    m = c5_latency_matrix(64)
    assert m.shape == (64, 64) and m.min() == 1 and m.max() == 5
    assert (np.diag(m) == 1).all() and not (m == m.T).all()          # asymmetric
    assert m[3][7] == 1 + (7 * 3 + 13 * 7) % 5
    assert bridge_ids(4, 2, 1024) == [0, 1, 128, 129, 256, 257, 384, 385]


def test_all_ones_matrix_is_the_plain_model(hostemu_lib):
    """Same seed, same script: a pool with an all-ones latency matrix and a pool with none end
    in identical state (columns, counters, digest), and so does a depth-2 pool column-wise."""
    n = 3000
    runs = []
    for depth, lat in ((2, None), (8, None), (8, np.ones((16, 16), dtype=np.uint8))):
        p = Pool(wan_config(hostemu_lib, capacity=n + 2, n_initial=n, seed=77, mailbox_depth=depth,
                            packet_loss_ppm=100000), hostemu_lib)
        p.latency_set(lat)
        x = p.member_add()
        p.join(x, [5])
        p.user_event(9, b"e", b"p", False)
        p.crash_many([20, 21])
        p.step(300)
        runs.append(p)
    sc.compare_pools(runs[1], runs[2], "ones vs none")                # includes the digest
    from parity import compare_columns, compare_stats
    compare_stats(runs[0], runs[1], "depth 2 vs 8")
    compare_columns(runs[0], runs[1], "depth 2 vs 8")


def test_latency_matrix_validation(hostemu_lib):
    p = Pool(wan_config(hostemu_lib, capacity=512, n_initial=512, seed=1), hostemu_lib)   # depth 2
    with pytest.raises(Exception):
        p.latency_set(np.full((2, 2), 2, dtype=np.uint8))             # needs a deeper ring
    p.latency_set(np.ones((2, 2), dtype=np.uint8))
    p8 = Pool(wan_config(hostemu_lib, capacity=512, n_initial=512, seed=1, mailbox_depth=8), hostemu_lib)
    with pytest.raises(Exception):
        p8.latency_set(np.full((2, 2), 8, dtype=np.uint8))
    with pytest.raises(Exception):
        p8.latency_set(np.zeros((2, 2), dtype=np.uint8))
    p8.latency_set(np.full((2, 2), 7, dtype=np.uint8))
    for bad in (3, 16):
        with pytest.raises(Exception):
            Pool(wan_config(hostemu_lib, capacity=8, n_initial=8, mailbox_depth=bad), hostemu_lib)


@pytest.mark.parametrize("cfg_fn,n_dcs", [(wan_config, 64), (lan_config, 5)])
def test_event_dissemination_parity(make, hostemu_lib, cfg_fn, n_dcs):
    """One user event over the C5 matrix: kernel body == oracle after every few ticks."""
    n = 64 * 128 + 77
    cfg = cfg_fn(hostemu_lib, capacity=n, n_initial=n, seed=0x5EED0005, mailbox_depth=8)
    pools = make(cfg)
    lat = c5_latency_matrix(n_dcs)
    for p in pools:
        p.latency_set(lat)
    slot = sc.both(pools, lambda p: p.user_event(0, b"deploy", b"x" * 32, False))
    sc.step_compare(pools, 12, 1, "early")
    t = sc.both(pools, lambda p: p.run_until(PRED_RUMOR_CONVERGED, slot, 800, 4))
    assert t != NEVER
    compare_pools(*pools, "converged")
    sc.step_compare(pools, 120, 30, "drain")
    for p in pools:
        s = p.stats()
        assert s["rumors_accepted"] == n - 1 and s["suspects"] == 0 and s["refutes"] == 0
        assert p.column("ltime_event")[:n].min() >= 2
    # latency only ever delays: the same pool without the matrix converges sooner
    q = Pool(cfg, hostemu_lib)
    s2 = q.user_event(0, b"deploy", b"x" * 32, False)
    t_plain = q.run_until(PRED_RUMOR_CONVERGED, s2, 800, 4)
    assert t_plain < t


def test_slow_acks_use_the_indirect_stage(make, hostemu_lib):
    """WAN timing (ProbeTimeout 6 ticks) with round trips of up to 8 extra ticks: slow direct acks
    go through the indirect/TCP stage and still succeed before the deadline — no suspicion."""
    n = 64 * 128
    cfg = wan_config(hostemu_lib, capacity=n, n_initial=n, seed=3, mailbox_depth=8)
    pools = make(cfg)
    for p in pools:
        p.latency_set(slow_matrix(64, 5))                              # round trips of 0..8 extra ticks
    sc.step_compare(pools, 60, 20, "steady")
    for p in pools:
        s = p.stats()
        assert s["probes"] > 0 and s["indirect_pings"] > 0            # some acks were late
        assert s["acks"] + (s["probes"] - s["acks"]) == s["probes"]
        assert s["suspects"] == 0 and s["probe_failures"] == 0
    # a plain pool never needs the indirect stage without loss
    q = Pool(wan_config(hostemu_lib, capacity=n, n_initial=n, seed=3), hostemu_lib)
    q.step(60)
    assert q.stats()["indirect_pings"] == 0


def test_lossy_crash_parity_with_latency(make, hostemu_lib):
    """Loss + crashes + no TCP fallback over the matrix: late acks, nacks, budgets, refutes."""
    n = 2048
    for tcp_off in (0, 1):
        cfg = wan_config(hostemu_lib, capacity=n + 2, n_initial=n, seed=21 + tcp_off, mailbox_depth=8,
                         packet_loss_ppm=200000, disable_tcp_pings=tcp_off)
        pools = make(cfg)
        for p in pools:
            p.latency_set(slow_matrix(16, 7))                          # some round trips miss every deadline
        x = sc.both(pools, lambda p: p.member_add())
        sc.both(pools, lambda p: p.join(x, [3]))
        sc.both(pools, lambda p: p.user_event(5, b"e1", b"payload", False))
        for p in pools:
            p.crash_many([10, 300, 1200])
        sc.step_compare(pools, 500, 25, f"lossy tcp_off={tcp_off}")
        s = pools[0].stats()
        assert s["packets_lost"] > 0 and s["nacks"] > 0 and s["suspects"] >= 3
        if tcp_off:
            assert s["refutes"] > 0
    td = sc.both(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 6000, 50))
    assert td != NEVER
    compare_pools(*pools, "crashed all dead")


def test_snapshot_restore_with_packets_in_flight(hostemu_lib):
    n = 4096
    cfg = wan_config(hostemu_lib, capacity=n, n_initial=n, seed=9, mailbox_depth=8)
    p = Pool(cfg, hostemu_lib)
    p.latency_set(c5_latency_matrix(32))
    slot = p.user_event(1, b"a", b"b", False)
    p.step(9)
    blob = p.snapshot()
    p.step(40)
    h1, info1 = p.state_hash(), p.rumor_info(slot)
    q = Pool(cfg, hostemu_lib)
    q.restore(blob)
    q.step(40)
    assert q.state_hash() == h1 and q.rumor_info(slot) == info1
    with pytest.raises(Exception):
        Pool(wan_config(hostemu_lib, capacity=n, n_initial=n, seed=9), hostemu_lib).restore(blob)


def _federation(make_pool, n, n_dcs, bridges, seed):
    pools = [make_pool(seed + k) for k in range(2)]
    fed = WanFederation(pools[0], pools[1], n_dcs=n_dcs, bridges_per_dc=bridges, n_members=n)
    fed.fire(0, 1, b"deploy", b"v2")
    return fed


def test_c5_two_pool_federation_parity(hostemu_lib):
    """BASELINE config 5 in small: two WAN pools, 16 datacenters, 3 bridges each; an event fired
    in A/DC0 reaches every member of A and B.  Kernel body and oracle agree after every tick."""
    n, n_dcs, bridges = 16 * 128 * 2, 16, 3
    cfg = lambda seed: wan_config(hostemu_lib, capacity=n, n_initial=n, seed=seed, mailbox_depth=8)
    fk = _federation(lambda s: Pool(cfg(s), hostemu_lib), n, n_dcs, bridges, 100)
    fo = _federation(lambda s: OraclePool(cfg(s)), n, n_dcs, bridges, 100)
    key = (b"deploy", b"v2")
    for tick in range(400):
        if fk.converged(key):
            break
        fk.step(1)
        fo.step(1)
        assert fk.slots == fo.slots and fk.forwarded == fo.forwarded
        if tick % 8 == 0:
            for a, b in zip(fk.pools, fo.pools):
                compare_pools(a, b, f"federation tick {tick}", columns=False)
    assert fk.converged(key) and fo.converged(key)
    for a, b in zip(fk.pools, fo.pools):
        compare_pools(a, b, "federation converged")
    assert fk.forwarded >= 1
    for x, (p, s) in enumerate(zip(fk.pools, fk.slots[key])):
        info = p.rumor_info(s)
        assert info["heard_count"] == n
        # exactly-once delivery: the origin, the bridge re-fires, and gossip add up to n
        assert p.stats()["rumors_accepted"] == n - (1 if x == 0 else 0) - fk.forwarded_into[x]
    # B was reached through the bridges only: its first holder is a bridge member
    assert fk.pools[1].rumor_info(fk.slots[key][1])["origin"] in fk.bridges
    # the crossing costs time: B converges no earlier than A's first bridge delivery + 1
    ta = fk.pools[0].rumor_info(fk.slots[key][0])["converged_tick"]
    tb = fk.pools[1].rumor_info(fk.slots[key][1])["converged_tick"]
    assert tb >= fk.pools[1].rumor_info(fk.slots[key][1])["start_tick"] > 0 and ta > 0
