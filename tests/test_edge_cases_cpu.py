"""Edge cases of the reference-facing surface, on the kernel body and the oracle side by side: empty
and one-member pools, exhaustion of the 30 tracked-broadcast slots, event-log overflow, joins that
reach nobody, repeated Leave / crash / force-leave, zero-length steps and horizons."""
import pytest

import scenarios as sc
from consul_b200.pool import (NEVER, PRED_ALL_RUMORS_CONVERGED, PRED_CRASHED_ALL_DEAD, PRED_RUMOR_CONVERGED,
                              GsimError, Pool, consul_test_config, lan_config)
from oracle_binding import OraclePool
from parity import compare_pools


@pytest.fixture()
def make(hostemu_lib):
    return lambda cfg: [Pool(cfg, hostemu_lib), OraclePool(cfg)]


def raises_both(pools, fn, code=None):
    for p in pools:
        with pytest.raises(Exception) as e:
            fn(p)
        if code is not None and isinstance(e.value, GsimError):
            assert e.value.code == code


def test_empty_pool_and_single_member(make, hostemu_lib):
    pools = make(consul_test_config(hostemu_lib, capacity=4, n_initial=0, seed=1, phase_group=1))
    sc.step_compare(pools, 10, 5, "empty")
    assert sc.both(pools, lambda p: p.stats()["n_members"]) == 0
    raises_both(pools, lambda p: p.members(0))
    a = sc.both(pools, lambda p: p.member_add(watched=True))
    assert a == 0
    sc.step_compare(pools, 30, 1, "alone")                      # probes nobody, gossips to nobody
    for p in pools:
        s = p.stats()
        assert s["probes"] == 0 and s["gossip_packets"] == 0 and p.members(0) == [(0, 1, 1, 0)]
    slot = sc.both(pools, lambda p: p.user_event(0, b"solo", b"", False))
    assert sc.both(pools, lambda p: p.run_until(PRED_RUMOR_CONVERGED, slot, 10, 1)) == pools[0].now  # at once
    assert sc.both(pools, lambda p: p.join(0, [0])) == 0         # joining yourself contacts nobody
    assert sc.both(pools, lambda p: p.join(0, [])) == 0
    compare_pools(*pools, "single")


def test_thirty_one_concurrent_events(make, hostemu_lib):
    """30 tracked broadcasts at once; the 31st is refused while all are still in flight and
    accepted once finished intents/alive messages can be folded away."""
    n = 200
    pools = make(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=3))
    slots = [sc.both(pools, lambda p: p.user_event(k, b"e%d" % k, b"x", False)) for k in range(30)]
    assert sorted(slots) == list(range(30))
    raises_both(pools, lambda p: p.user_event(31, b"one too many", b"", False), code=-4)
    sc.step_compare(pools, 100, 20, "30 rumors")
    for p in pools:
        assert all(p.rumor_info(s)["heard_count"] == n for s in slots)
        assert p.stats()["rumors_accepted"] == 30 * (n - 1)
    raises_both(pools, lambda p: p.user_event(31, b"still full", b"", False), code=-4)   # events are kept
    for p in pools:
        p.rumor_retire(slots[7])
    assert sc.both(pools, lambda p: p.user_event(31, b"fits now", b"", False)) == slots[7]
    sc.step_compare(pools, 60, 20, "after retire")


def test_event_log_overflow_is_counted(make, hostemu_lib):
    n = 300
    pools = make(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=4, flags=1, event_log_capacity=16))
    for p in pools:
        for w in range(0, n, 3):
            p.member_watch(w, True)
    sc.both(pools, lambda p: p.user_event(1, b"e", b"", False))
    for p in pools:
        p.step(60)
    ev = [p.poll_events() for p in pools]
    assert len(ev[0]) == 16 and len(ev[1]) == 16                 # the ring kept its capacity ...
    for p in pools:
        assert p.stats()["events_dropped"] == 100 - 16           # ... and counted the rest (100 watchers)
    compare_pools(*pools, "after overflow", columns=True)


def test_join_that_reaches_nobody_and_repeated_operations(make, hostemu_lib):
    n = 64
    pools = make(lan_config(hostemu_lib, capacity=n + 4, n_initial=n, seed=5, flags=1))
    for p in pools:
        p.crash(9)
        p.crash(9)                                               # crashing a crashed member: no-op
    x = sc.both(pools, lambda p: p.member_add())
    assert sc.both(pools, lambda p: p.join(x, [9])) == 0          # the only seed is down
    assert sc.both(pools, lambda p: p.num_nodes(x)) == 1          # still alone
    assert sc.both(pools, lambda p: p.join(x, [9, 10, n + 50])) == 1   # one live seed among bad ones
    for p in pools:
        p.leave(3)
    raises_both(pools, lambda p: p.leave(3), code=-6)             # already leaving
    raises_both(pools, lambda p: p.leave(9), code=-6)             # not running
    for p in pools:
        p.force_leave(0, 20)                                      # alive member: nothing to remove
        assert dict((m[0], m[1]) for m in p.members(0))[20] == 1
    sc.step_compare(pools, 400, 50, "aftermath")
    for p in pools:
        p.force_leave(0, 9, True)                                 # Failed -> Left -> erased
        p.force_leave(0, 9, True)                                 # again: no-op
        assert 9 not in [m[0] for m in p.members(0)]
    compare_pools(*pools, "end")


def test_zero_length_horizons(make, hostemu_lib):
    pools = make(lan_config(hostemu_lib, capacity=32, n_initial=32, seed=6))
    for p in pools:
        p.step(0)
        assert p.now == 0
        assert p.run_until(PRED_CRASHED_ALL_DEAD, 0, 0, 1) == NEVER           # nothing crashed, no ticks
        assert p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 0, 1) == 0           # no rumors: vacuously now
        assert p.now == 0
    raises_both(pools, lambda p: p.run_until(PRED_RUMOR_CONVERGED, 0, 10, 0))  # check_every = 0
    raises_both(pools, lambda p: p.run_until(99, 0, 10, 1))                    # unknown predicate
    raises_both(pools, lambda p: p.rumor_info(5))                              # free slot
    compare_pools(*pools, "untouched")


def test_identical_event_fired_by_a_second_member(make, hostemu_lib):
    """[U] serf.UserEvent: the (LTime, Name, Payload) de-dup is per member buffer — a second member that
    fires the identical event while it has not seen the first one delivers it itself (exactly once)
    and queues its own broadcast; a member that already holds it re-queues with transmits = 0."""
    n = 400
    pools = make(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=11))
    for p in pools:
        p.member_watch(7, True)
    s0 = sc.both(pools, lambda p: p.user_event(3, b"deploy", b"v1", False))
    s1 = sc.both(pools, lambda p: p.user_event(7, b"deploy", b"v1", False))     # same LTime 1, same bytes
    assert s1 == s0
    for p in pools:
        assert p.rumor_info(s0)["heard_count"] == 2
        assert [(e.type, e.observer) for e in p.poll_events()] == [(5, 7)]       # EventUser at member 7, once
    s2 = sc.both(pools, lambda p: p.user_event(3, b"deploy", b"v1", False))     # origin again: LTime now 2 -> new event
    assert s2 != s0
    sc.step_compare(pools, 80, 10, "two origins")
    for p in pools:
        assert p.rumor_info(s0)["heard_count"] == n and p.rumor_info(s2)["heard_count"] == n
    compare_pools(*pools, "dedup")


def test_encoded_user_event_size_limit(make, hostemu_lib):
    """name+payload may pass the first check and still be refused once encoded (second check of
    [U] serf.UserEvent): 500 raw bytes encode to more than UserEventSizeLimit = 512."""
    pools = make(lan_config(hostemu_lib, capacity=8, n_initial=8, seed=2))
    raises_both(pools, lambda p: p.user_event(0, b"n" * 250, b"p" * 250, False), code=-7)
    assert sc.both(pools, lambda p: p.user_event(0, b"n" * 200, b"p" * 200, False)) is not None
