"""Scenario scripts replayed on two pools and compared after every phase.

Each scenario mirrors a reference test or a BASELINE config:
  c1  TestServer_JoinLAN / TestServer_LANReap shape  (agent/consul/server_test.go:509-529, 666-733)
  c2  BASELINE config 2: single join cascade into a converged pool
  c3  BASELINE config 3: crash injection, suspicion -> dead convergence
  c4  BASELINE config 4 / TestClientServer_UserEvent (client_test.go:756-835): one user event
"""
from __future__ import annotations

from consul_b200.pool import (NEVER, PRED_ALL_RUMORS_CONVERGED, PRED_CRASHED_ALL_DEAD,
                              PRED_RUMOR_CONVERGED, consul_test_config, lan_config, wan_config)
from parity import compare_pools

STATUS_ALIVE, STATUS_LEFT, STATUS_FAILED = 1, 3, 4


def both(pools, fn):
    out = [fn(p) for p in pools]
    assert out[0] == out[1], (out[0], out[1])
    return out[0]


def step_compare(pools, ticks, every, where, columns=True):
    done = 0
    while done < ticks:
        k = min(every, ticks - done)
        for p in pools:
            p.step(k)
        done += k
        compare_pools(pools[0], pools[1], f"{where} +{done}", columns=columns)


def c1_three_node_join(make, lib, seed=1):
    """3 agents, s2 and s3 join s1 (server_test.go:704-705), then s3 crashes (":725")."""
    cfg = consul_test_config(lib, capacity=8, n_initial=0, seed=seed, flags=1, phase_group=1)
    pools = make(cfg)
    ids = [both(pools, lambda p: p.member_add(watched=True)) for _ in range(3)]
    assert ids == [0, 1, 2]
    for p in pools:
        assert [m[0] for m in p.members(0)] == [0]          # alone before any join
    assert both(pools, lambda p: p.join(1, [0])) == 1
    assert both(pools, lambda p: p.join(2, [0])) == 1
    compare_pools(*pools, "after joins")
    t = both(pools, lambda p: p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 400, 1))
    assert t != NEVER
    for p in pools:
        for obs in ids:
            assert sorted((m[0], m[1]) for m in p.members(obs)) == [(i, STATUS_ALIVE) for i in ids]
    ev = both(pools, lambda p: [(e.tick, e.type, e.subject, e.observer) for e in p.poll_events()])
    joins = {(e[3], e[2]) for e in ev if e[1] == 0}
    assert {(1, 0), (0, 1), (2, 0), (0, 2), (1, 2), (2, 1)} <= joins   # everyone saw everyone join
    step_compare(pools, 40, 1, "settle")
    # crash s3 without Leave(): survivors must see it Failed (LANReap before the reap)
    for p in pools:
        p.crash(2)
    td = both(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 2000, 1))
    assert td != NEVER
    compare_pools(*pools, "after failure")
    for p in pools:
        assert dict((m[0], m[1]) for m in p.members(0))[2] == STATUS_FAILED
        s = p.stats()
        assert s["refutes"] == 0 and s["n_view_dead"] == 1
    # force-leave turns Failed into Left (agent_endpoint_test.go:2524-2566)
    for p in pools:
        p.force_leave(0, 2)
        assert dict((m[0], m[1]) for m in p.members(1))[2] == STATUS_LEFT
    compare_pools(*pools, "after force-leave")
    return t, td


def c2_join_cascade(make, lib, n, seed=0x5EED0001, extra_ticks=64, every=8, columns=True):
    cfg = lan_config(lib, capacity=n + 1, n_initial=n, seed=seed)
    pools = make(cfg)
    x = both(pools, lambda p: p.member_add())
    assert both(pools, lambda p: p.join(x, [0])) == 1
    t = both(pools, lambda p: p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 600, every))
    assert t != NEVER
    compare_pools(*pools, "converged", columns=columns)
    step_compare(pools, extra_ticks, extra_ticks, "steady", columns=columns)
    for p in pools:
        s = p.stats()
        assert s["n_up"] == n + 1 and s["suspects"] == 0 and s["probe_failures"] == 0
        # join intent reaches all n by gossip; alive reaches n-1 (the seed got it in the push-pull)
        assert s["rumors_accepted"] == 2 * n - 1
        assert len(p.members(x)) == n + 1
    return t


def c3_crash(make, lib, n, ppm=100000, seed=0x5EED0002, every=50, columns=True, cfg_fn=lan_config, **kw):
    cfg = cfg_fn(lib, capacity=n, n_initial=n, seed=seed, **kw)
    pools = make(cfg)
    crashed = both(pools, lambda p: p.crash_fraction(ppm, 3))
    assert crashed > 0
    st = pools[0].stats()
    horizon = st["suspicion_ticks"][0] + 4 * st["probe_interval_ticks"] * 8 + 200
    t = both(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, horizon + 4000, every))
    assert t != NEVER
    compare_pools(*pools, "all dead", columns=columns)
    for p in pools:
        s = p.stats()
        assert s["n_crashed"] == crashed and s["n_view_dead"] == crashed
        assert s["deads"] == crashed and s["refutes"] == 0       # lossless: no false positives
        assert s["suspects"] == crashed
    return crashed, t


def c4_user_event(make, lib, n, seed=0x5EED0003, columns=True):
    cfg = lan_config(lib, capacity=n, n_initial=n, seed=seed)
    pools = make(cfg)
    slot = both(pools, lambda p: p.user_event(0, b"deploy", b"x" * 32, False))
    t = both(pools, lambda p: p.run_until(PRED_RUMOR_CONVERGED, slot, 600, 4))
    assert t != NEVER
    compare_pools(*pools, "event delivered", columns=columns)
    for p in pools:
        info = p.rumor_info(slot)
        assert info["heard_count"] == n and info["ltime"] == 1
        lt = p.column("ltime_event")[:n]
        assert lt.min() >= 2                                   # everyone witnessed LTime 1
        s = p.stats()
        assert s["rumors_accepted"] == n - 1                   # exactly-once delivery
    step_compare(pools, 80, 80, "drain", columns=columns)
    for p in pools:
        assert p.rumor_info(slot)["queued_count"] == 0
    return t


def leave_scenario(make, lib, n=200, seed=5):
    cfg = lan_config(lib, capacity=n, n_initial=n, seed=seed, flags=1)
    pools = make(cfg)
    for p in pools:
        p.leave(7)
        assert dict((m[0], m[1]) for m in p.members(0))[7] == STATUS_LEFT
    compare_pools(*pools, "leaving")
    step_compare(pools, 120, 10, "leave drain")
    for p in pools:
        s = p.stats()
        assert s["n_gone"] == 1 and s["n_up"] == n - 1 and s["refutes"] == 0
        assert s["suspects"] == 0                              # a clean leave is never suspected
    ev = both(pools, lambda p: [(e.tick, e.type, e.subject) for e in p.poll_events()])
    assert (0, 1, 7) in ev                                     # EventMemberLeave


def lossy_scenario(make, lib, n=600, loss_ppm=150000, seed=11, ticks=600, cfg_fn=lan_config, **kw):
    """Packet loss exercises indirect probes, nacks, awareness, false suspicion and refutation."""
    cfg = cfg_fn(lib, capacity=n + 2, n_initial=n, seed=seed, packet_loss_ppm=loss_ppm, **kw)
    pools = make(cfg)
    x = both(pools, lambda p: p.member_add())
    both(pools, lambda p: p.join(x, [3]))
    both(pools, lambda p: p.user_event(5, b"e1", b"payload", False))
    for p in pools:
        p.crash_many([10, 11, 12])
    step_compare(pools, ticks, 20, "lossy")
    s = pools[0].stats()
    assert s["packets_lost"] > 0 and s["indirect_pings"] > 0
    return s


def budget_scenario(make, lib, n=300, seed=21):
    """A 120-byte UDP budget forces TransmitLimitedQueue ordering (tiers, sizes, classes)."""
    cfg = lan_config(lib, capacity=n + 8, n_initial=n, seed=seed, udp_buffer_size=120)
    pools = make(cfg)
    for k in range(4):
        x = both(pools, lambda p: p.member_add(alive_msg_size=30 + 7 * k))
        both(pools, lambda p: p.join(x, [k]))
        both(pools, lambda p: p.user_event(20 + k, b"ev%d" % k, b"p" * (5 + 9 * k), False))
        step_compare(pools, 3, 1, f"budget round {k}")
    step_compare(pools, 150, 5, "budget drain")


def event_window_scenario(make, lib, n=64, seed=31):
    """ignore_old joins set eventMinTime; the 512-entry window drops ancient events."""
    cfg = lan_config(lib, capacity=n + 4, n_initial=n, seed=seed, event_buffer=4)
    pools = make(cfg)
    s0 = both(pools, lambda p: p.user_event(1, b"old", b"1", False))
    step_compare(pools, 4, 1, "first event")
    x = both(pools, lambda p: p.member_add(watched=True))
    both(pools, lambda p: p.join(x, [2], True))                 # ignore_old: must not replay "old"
    for p in pools:
        assert not (int(p.column("heard")[x]) >> s0) & 1
    # push the clock of member 1 far ahead, then fire an event: members with a far-ahead
    # clock drop LTime values older than clock - event_buffer
    for k in range(6):
        both(pools, lambda p: p.user_event(1, b"burst%d" % k, b"", False))
    step_compare(pools, 60, 3, "window")
    y = both(pools, lambda p: p.member_add(watched=True))
    both(pools, lambda p: p.join(y, [1], False))                # replay: only recent ones accepted
    compare_pools(*pools, "replay join")
    step_compare(pools, 40, 5, "after replay")
    return pools


def lan_reap_scenario(make, lib, seed=1):
    """TestServer_LANReap (agent/consul/server_test.go:666-733): ReconnectTimeout = TombstoneTimeout
    = 250 ms, ReapInterval = 300 ms; three servers converge, s2 shuts down without leaving, and
    the survivors' member lists shrink from 3 to 2 once it has been Failed and then reaped."""
    MS = 1_000_000
    cfg = consul_test_config(lib, capacity=8, n_initial=0, seed=seed, flags=1, phase_group=1,
                             reconnect_timeout_ns=250 * MS, tombstone_timeout_ns=250 * MS,
                             reap_interval_ns=300 * MS)
    pools = make(cfg)
    ids = [both(pools, lambda p: p.member_add(watched=True)) for _ in range(3)]
    both(pools, lambda p: p.join(1, [0]))
    both(pools, lambda p: p.join(2, [0]))
    t = both(pools, lambda p: p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 400, 1))
    assert t != NEVER
    for p in pools:
        for obs in ids:
            assert len(p.members(obs)) == 3
    for p in pools:
        p.crash(1)                                            # s2.Shutdown(): no Leave
    td = both(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 2000, 1))
    assert td != NEVER
    for p in pools:
        assert dict((m[0], m[1]) for m in p.members(0))[1] == STATUS_FAILED
    # Failed for > 5 ticks, then the next reaper wake-up (every 6 ticks) erases it
    seen = None
    for k in range(20):
        step_compare(pools, 1, 1, f"reap wait {k}")
        if len(pools[0].members(0)) == 2:
            seen = pools[0].now
            break
    assert seen is not None and seen - td > 5 and seen - td <= 5 + 6 + 1
    for p in pools:
        for obs in (0, 2):
            assert sorted(m[0] for m in p.members(obs)) == [0, 2]
        s = p.stats()
        assert s["n_members"] == 3 and s["n_crashed"] == 0 and s["n_view_dead"] == 0
    ev = both(pools, lambda p: [(e.type, e.subject) for e in p.poll_events()])
    assert (2, 1) in ev and (4, 1) in ev                       # EventMemberFailed, then EventMemberReap
    assert ev.index((2, 1)) < ev.index((4, 1))
    # a clean leave is reaped after TombstoneTimeout as well
    for p in pools:
        p.leave(2)
    step_compare(pools, 120, 4, "leave + tombstone")
    for p in pools:
        assert sorted(m[0] for m in p.members(0)) == [0]
    return td, seen


def set_tags_scenario(make, lib, n=500, seed=41):
    """(*Serf).SetTags -> memberlist.UpdateNode: next incarnation, alive re-broadcast, one
    EventMemberUpdate per other member; a probe that was in flight against the old incarnation
    cannot hurt the member any more."""
    cfg = lan_config(lib, capacity=n + 2, n_initial=n, seed=seed, packet_loss_ppm=200000)
    pools = make(cfg)
    for p in pools:
        for w in (3, 77, n - 1):
            p.member_watch(w, True)
    step_compare(pools, 7, 1, "before")
    slot = both(pools, lambda p: p.member_update(5, 120))
    for p in pools:
        assert int(p.column("key")[5]) >> 5 == 2              # nextIncarnation
        info = p.rumor_info(slot)
        assert info["kind"] == 5 and info["subject"] == 5 and info["incarnation"] == 2
    t = both(pools, lambda p: p.run_until(PRED_RUMOR_CONVERGED, slot, 600, 2))
    assert t != NEVER
    step_compare(pools, 100, 10, "drain")
    ev = both(pools, lambda p: sorted((e.type, e.subject, e.observer) for e in p.poll_events()))
    assert [e for e in ev if e[0] == 3] == [(3, 5, 3), (3, 5, 77), (3, 5, n - 1)]   # EventMemberUpdate once each
    for p in pools:
        assert len(p.members(5)) == n
    # a second update before the first retires, and an update by a leaving member is refused
    s2 = both(pools, lambda p: p.member_update(5))
    assert s2 != slot or True
    for p in pools:
        assert int(p.column("key")[5]) >> 5 == 3
        p.leave(9)
        try:
            p.member_update(9)
            raise AssertionError("update of a leaving member must fail")
        except AssertionError:
            raise
        except Exception:
            pass
    step_compare(pools, 60, 10, "after second update")


def short_reconnect_timeout_scenario(make, lib, seed=1):
    """TestClient_ShortReconnectTimeout (agent/consul/client_test.go:862-894): ReapInterval 50 ms
    everywhere, the clients advertise a reconnect timeout of 100 ms (libserf/serf.go:68-85 turns
    the tag into serf's ReconnectTimeoutOverride); the pool default stays at 72 h.  A client that
    shuts down is forgotten by the others shortly after it is declared Failed; a failed SERVER
    (no override) stays Failed."""
    MS = 1_000_000
    cfg = consul_test_config(lib, capacity=8, n_initial=0, seed=seed, flags=1, phase_group=1, reap_interval_ns=50 * MS)
    pools = make(cfg)
    server, server2, c0, c1 = [both(pools, lambda p: p.member_add(watched=True)) for _ in range(4)]
    for p in pools:
        p.member_reconnect_timeout_set(c0, 100 * MS)
        p.member_reconnect_timeout_set(c1, 100 * MS)
    for m in (server2, c0, c1):
        both(pools, lambda p: p.join(m, [server]))
    assert both(pools, lambda p: p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 400, 1)) != NEVER
    for p in pools:
        assert len(p.members(server)) == 4
        p.crash(c1)
    td = both(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 2000, 1))
    assert td != NEVER
    step_compare(pools, 4, 1, "reap the client")
    for p in pools:
        assert sorted(m[0] for m in p.members(server)) == [server, server2, c0]
        assert sorted(m[0] for m in p.members(c0)) == [server, server2, c0]
        p.crash(server2)
    td2 = both(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 2000, 1))
    assert td2 != NEVER
    step_compare(pools, 200, 20, "a server without override stays Failed")
    for p in pools:
        assert dict((m[0], m[1]) for m in p.members(server))[server2] == STATUS_FAILED
        p.member_reconnect_timeout_set(server2, 100 * MS)        # tags changed: now it may be forgotten
    step_compare(pools, 3, 1, "override set later")
    for p in pools:
        assert sorted(m[0] for m in p.members(server)) == [server, c0]
    ev = both(pools, lambda p: [(e.type, e.subject) for e in p.poll_events()])
    assert (4, c1) in ev and (4, server2) in ev                   # EventMemberReap for both
