"""BASELINE's full-size configurations against the ORACLE (not only through properties): the CUDA
path and oracle/liboracle.so run the same script at 4 000 000 (C3), 16 777 216 (C4) and
2 x 8 388 608 (C5) members; the 256-bit state digest and every counter are compared at checkpoints
every few hundred ticks, plus the exact convergence ticks.  (VERDICT r1, weak #2.)  The oracle uses
every host thread; the GPU side is the sm_100a path through the C ABI."""
import pytest

from consul_b200.pool import (NEVER, PRED_CRASHED_ALL_DEAD, PRED_RUMOR_CONVERGED, Pool, lan_config, wan_config)
from consul_b200.wan import WanFederation
from oracle_binding import OraclePool
from parity import compare_pools

pytestmark = pytest.mark.gpu


def both(pools, fn):
    a, b = [fn(p) for p in pools]
    assert a == b, (a, b)
    return a


def test_c3_4m_crash_wave_against_the_oracle(cuda_lib):
    """C3: 4 000 000 members, ~10 % crashed at tick 0, 2 000 ticks: first suspicions, the Lifeguard
    timers with confirmations, every crashed member Dead — digest and counters equal throughout."""
    n = 4_000_000
    cfg = lan_config(cuda_lib, capacity=n, n_initial=n, seed=0x5EED0003)
    pools = [Pool(cfg, cuda_lib), OraclePool(cfg, threads=0)]
    crashed = both(pools, lambda p: p.crash_fraction(100000, 0))
    assert abs(crashed - n // 10) < 5000
    for upto in (16, 64, 200, 300, 500, 900, 1400, 2000):
        for p in pools:
            p.step(upto - p.now)
        compare_pools(*pools, f"C3 tick {upto}", columns=False)
    t_dead = both(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 0, 1))   # already true: the recorded tick
    assert t_dead != NEVER and 265 <= t_dead <= 2000
    s = pools[0].stats()
    assert s["deads"] == crashed and s["refutes"] == 0 and s["n_view_suspect"] == 0


def test_c4_16m_user_event_against_the_oracle(cuda_lib):
    """C4 on one GPU: 16 777 216 members, one user event, until converged and drained."""
    n = 16_777_216
    cfg = lan_config(cuda_lib, capacity=n, n_initial=n, seed=0x5EED0004)
    pools = [Pool(cfg, cuda_lib), OraclePool(cfg, threads=0)]
    slot = both(pools, lambda p: p.user_event(0, b"deploy", bytes(32), False))
    for upto in (8, 24, 40, 72, 136):
        for p in pools:
            p.step(upto - p.now)
        compare_pools(*pools, f"C4 tick {upto}", columns=False)
    info = both(pools, lambda p: p.rumor_info(slot))
    assert info["heard_count"] == n and 20 < info["converged_tick"] < 80
    s = pools[0].stats()
    assert s["rumors_accepted"] == n - 1 and s["rumors_sent"] == 32 * n      # exactly once; budget 4*ceil(log10(n+1))


def test_c5_2x8m_wan_federation_against_the_oracle(cuda_lib):
    """C5: two WAN pools of 8 388 608 members, 64 datacenters, the asymmetric latency matrix, 5 bridges
    per datacenter; the event crosses from A to B through the bridges.  Same driver on both sides."""
    n = 8 * 1024 * 1024
    cfg = lambda seed: wan_config(cuda_lib, capacity=n, n_initial=n, seed=seed, mailbox_depth=8)
    feds = [WanFederation(Pool(cfg(0x5EED0051), cuda_lib), Pool(cfg(0x5EED0052), cuda_lib), 64, 5, n),
            WanFederation(OraclePool(cfg(0x5EED0051), threads=0), OraclePool(cfg(0x5EED0052), threads=0), 64, 5, n)]
    for f in feds:
        f.fire(0, 7, b"deploy", b"x" * 32)
    t = [f.run_until_converged(b"deploy", b"x" * 32, 400) for f in feds]
    assert t[0] == t[1] and t[0] is not None and 10 < t[0] < 200
    assert feds[0].forwarded == feds[1].forwarded >= 1 and feds[0].forwarded_into == feds[1].forwarded_into
    for f in feds:
        f.step(16)
    for x in (0, 1):
        compare_pools(feds[0].pools[x], feds[1].pools[x], f"C5 pool {'AB'[x]}", columns=False)
        assert feds[0].pools[x].rumor_info(feds[0].slots[(b"deploy", b"x" * 32)][x])["heard_count"] == n
