"""Scenario-level pins of the oracle against what the reference itself states or asserts
(SURVEY.md §8c "statistical cross-check" and the eventual outcomes of its tests)."""
import math

from consul_b200 import _lib
from consul_b200.pool import NEVER, PRED_RUMOR_CONVERGED, lan_config, wan_config
from oracle_binding import OraclePool


def test_leave_propagate_design_figure():
    """/root/reference/internal/gossip/libserf/serf.go:29-33: LeavePropagateDelay = 3 s was chosen
    so that a leave reaches > 99.99 % of a 100 000-node cluster (LAN defaults: 15 gossip rounds,
    fan-out 3).  A single rumor in the oracle must do at least that well."""
    n = 100_000
    L = _lib.lib()
    o = OraclePool(lan_config(L, capacity=n, n_initial=n, seed=0x5EED0001), threads=0)
    slot = o.user_event(0, b"x", b"", False)
    o.step(30)                                   # 3 s = 15 gossip rounds of 200 ms
    heard = o.rumor_info(slot)["heard_count"]
    assert heard / n > 0.9999, heard
    t = o.run_until(PRED_RUMOR_CONVERGED, slot, 200, 1)
    assert t != NEVER and t < 40


def test_dissemination_scales_logarithmically():
    """Epidemic spread with fan-out 3: convergence ticks grow ~ log(N), and every member sends
    each rumor exactly RetransmitMult * ceil(log10(N+1)) times (runtime.go:1328-1330)."""
    L = _lib.lib()
    ticks = []
    for n in (1_000, 10_000, 100_000):
        o = OraclePool(lan_config(L, capacity=n, n_initial=n, seed=7), threads=0)
        slot = o.user_event(0, b"e", b"p", False)
        t = o.run_until(PRED_RUMOR_CONVERGED, slot, 300, 1)
        o.step(100)
        s = o.stats()
        limit = 4 * math.ceil(math.log10(n + 1))
        assert s["retransmit_limit"] == limit and s["rumors_sent"] == limit * n
        assert s["rumors_accepted"] == n - 1
        ticks.append(t)
    assert ticks[0] < ticks[1] < ticks[2] <= ticks[0] + 16


def test_suspicion_bounds_lan_and_wan():
    """A crashed member is declared dead no earlier than the Lifeguard minimum and no later than
    the maximum + one probe pass (runtime.go:1310-1312: SuspicionMult * log(N+1) * ProbeInterval)."""
    L = _lib.lib()
    for cfg_fn, n in ((lan_config, 2000), (wan_config, 1000)):
        o = OraclePool(cfg_fn(L, capacity=n, n_initial=n, seed=3), threads=0)
        o.crash_many([5, 77, 400])
        st = o.stats()
        lo, hi = st["suspicion_ticks"][st["suspicion_k"]], st["suspicion_ticks"][0]
        t = o.run_until(3, 0, hi + 400, 1)
        assert t != NEVER and lo <= t <= hi + 30 * st["probe_interval_ticks"]
        s = o.stats()
        assert s["deads"] == 3 and s["refutes"] == 0 and s["n_view_dead"] == 3
