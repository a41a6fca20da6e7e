"""Cross-check of the PROJECTION the product and the oracle share ("M1": one shared record per
subject, exact per-member tracking of broadcasts) against "M0" (oracle/m0_memberlist.py): a
full-fidelity restatement of memberlist + serf's piggyback in which every agent keeps its own view
of every other agent, its own suspicion timers and real broadcast queues (SURVEY.md §7 hard part
2, §8d C1 "run on M0 and M1 oracles").

M0 and M1 use randomness differently, so nothing here is bit-for-bit.  What must agree:
  * counts the protocol fixes exactly (every agent sends a broadcast RetransmitMult*ceil(log10(n+1))
    times; an event is delivered exactly once; Lamport clocks after an event);
  * eventual outcomes (who is Alive / Failed / Left for whom; no false suspicion without loss);
  * times within the protocol's own bounds and with matching means over seeds (dissemination
    rounds, first Dead declaration under Lifeguard's confirmations).
"""
import os
import statistics as st
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import m0_memberlist as m0  # noqa: E402

from consul_b200 import _lib  # noqa: E402
from consul_b200.pool import (NEVER, PRED_ALL_RUMORS_CONVERGED, PRED_CRASHED_ALL_DEAD, PRED_RUMOR_CONVERGED,  # noqa: E402
                              Pool, consul_test_config, lan_config)
from oracle_binding import OraclePool  # noqa: E402


def m1_pool(cfg, hostemu_lib, which):
    return OraclePool(cfg) if which == "oracle" else Pool(cfg, hostemu_lib)


@pytest.mark.parametrize("which", ["oracle", "kernel-body"])
def test_event_dissemination_matches_full_fidelity_model(which, hostemu_lib):
    n, seeds = 200, 12
    limit = m0.retransmit_limit(4, n)
    t0, t1 = [], []
    for seed in range(seeds):
        net = m0.Network(m0.Config(), seed=seed)
        net.converged_cluster(n)
        key = net.user_event(0, b"deploy", b"x" * 32)
        t = net.first_tick(lambda: all(any(k == key for _, k in a.delivered) for a in net.agents), 400)
        assert t is not None
        net.step(150)
        assert net.stats["msgs"] == limit * n                          # every agent sent it `limit` times
        assert all(sum(1 for _, k in a.delivered if k == key) == 1 for a in net.agents)   # exactly once
        assert all(a.clock_event == 2 for a in net.agents)              # everyone witnessed LTime 1
        t0.append(t)

        p = m1_pool(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=1000 + seed, phase_group=1), hostemu_lib, which)
        slot = p.user_event(0, b"deploy", b"x" * 32, False)
        t = p.run_until(PRED_RUMOR_CONVERGED, slot, 400, 1)
        assert t != NEVER
        p.step(150)
        s = p.stats()
        assert s["rumors_sent"] == limit * n and s["rumors_accepted"] == n - 1
        assert int(p.column("ltime_event")[:n].min()) == 2 and int(p.column("ltime_event")[:n].max()) == 2
        t1.append(t)
    assert abs(st.mean(t0) - st.mean(t1)) <= 1.5, (t0, t1)               # same number of gossip rounds
    assert max(t0) <= 20 and max(t1) <= 20


@pytest.mark.parametrize("which", ["oracle", "kernel-body"])
def test_failure_detection_matches_full_fidelity_model(which, hostemu_lib):
    """Three crashed agents among 100, no loss.  M0: the first observer to declare each one Dead
    (after Lifeguard's confirmations shortened its timer) gossips it; M1's shared record stands
    for exactly that earliest timer."""
    n, crashed, seeds = 100, (3, 40, 77), 8
    first0, dead1 = [], []
    for seed in range(seeds):
        net = m0.Network(m0.Config(), seed=seed)
        net.converged_cluster(n)
        for c in crashed:
            net.crash(c)
        first = {}
        for _ in range(1500):
            net.step(1)
            for c in crashed:
                if c not in first and any(a.views[c].state == m0.DEAD for a in net.up_agents()):
                    first[c] = net.now
            if all(net.all_see(c, m0.DEAD) for c in crashed):
                break
        assert all(net.all_see(c, m0.DEAD) for c in crashed)            # everybody agrees in the end
        assert not any(stt.state != m0.ALIVE for a in net.up_agents() for x, stt in a.views.items() if x not in crashed)
        assert all(a.stats["refutes"] == 0 for a in net.agents)         # no false suspicion without loss
        first0 += list(first.values())

        p = m1_pool(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=2000 + seed, phase_group=1), hostemu_lib, which)
        p.crash_many(list(crashed))
        assert p.run_until(PRED_CRASHED_ALL_DEAD, 0, 3000, 1) != NEVER
        s = p.stats()
        assert s["deads"] == 3 and s["refutes"] == 0 and s["suspects"] == 3 and s["n_view_dead"] == 3
        ct = p.column("change_tick")
        dead1 += [int(ct[c]) for c in crashed]
        lo, hi = s["suspicion_ticks"][s["suspicion_k"]], s["suspicion_ticks"][0]
        for x in dead1[-3:] + list(first.values()):
            assert lo <= x <= hi + 3 * s["probe_interval_ticks"]         # both inside Lifeguard's [min, max]
    assert abs(st.mean(first0) - st.mean(dead1)) <= 6.0, (first0, dead1)


@pytest.mark.parametrize("which", ["oracle", "kernel-body"])
def test_three_node_join_crash_leave_outcomes(which, hostemu_lib):
    """BASELINE config 1 (TestServer_JoinLAN / LANReap shape, Consul's test timing) on both models."""
    cfg0 = m0.Config(probe_interval=2, probe_timeout=1, gossip_interval=2, suspicion_mult=2, tick_seconds=0.05)
    for seed in (1, 2, 3):
        net = m0.Network(cfg0, seed=seed)
        a, b, c = net.create(), net.create(), net.create()
        assert net.join(b, a) == 1 and net.join(c, a) == 1
        t = net.first_tick(lambda: all(len(x.views) == 3 for x in net.agents), 200)
        assert t is not None and t < 40
        net.crash(c)
        td = net.first_tick(lambda: net.all_see(c, m0.DEAD), 600)
        assert td is not None
        net.leave(b)
        tl = net.first_tick(lambda: net.state_of(a, b) == m0.LEFT, 200)
        assert tl is not None and net.agents[a].stats["refutes"] == 0

        p = m1_pool(consul_test_config(hostemu_lib, capacity=8, n_initial=0, seed=seed, phase_group=1), hostemu_lib, which)
        ids = [p.member_add() for _ in range(3)]
        p.join(ids[1], [ids[0]])
        p.join(ids[2], [ids[0]])
        t1 = p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 200, 1)
        assert t1 != NEVER and t1 < 40 and all(len(p.members(i)) == 3 for i in ids)
        p.crash(ids[2])
        td1 = p.run_until(PRED_CRASHED_ALL_DEAD, 0, 600, 1)
        assert td1 != NEVER
        assert abs((td1 - t1) - (td - t)) <= 12                             # same order of detection delay
        p.leave(ids[1])
        assert dict((m[0], m[1]) for m in p.members(ids[0]))[ids[1]] == 3    # Left at once
        assert p.stats()["refutes"] == 0


def test_m0_refutes_under_loss_like_m1(hostemu_lib):
    """40 % loss without the TCP fallback: live agents get suspected and refute with a higher
    incarnation in both models; nobody live ends up Dead for long."""
    n = 60
    net = m0.Network(m0.Config(loss=0.4, disable_tcp=True), seed=5)
    net.converged_cluster(n)
    net.step(600)
    refutes0 = sum(a.stats["refutes"] for a in net.agents)
    assert refutes0 > 0 and max(a.inc for a in net.agents) > 1
    p = OraclePool(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=5, phase_group=1, packet_loss_ppm=400000,
                              disable_tcp_pings=1))
    p.step(600)
    s = p.stats()
    assert s["refutes"] > 0 and int((p.column("key")[:n] >> 5).max()) > 1
    # same order of magnitude of false suspicion (per-observer suspicion in M0 vs one shared record in M1)
    assert 0.2 <= (s["refutes"] + 1) / (refutes0 + 1) <= 5.0, (s["refutes"], refutes0)


def test_join_cascade_matches_full_fidelity_model(hostemu_lib):
    """BASELINE config 2 in small: one joiner into a converged cluster of 300; ticks until everybody
    lists it.  M0: every agent's own memberlist has to learn the alive message; M1: the tracked
    alive rumor's heard-bits."""
    n, seeds = 300, 10
    t0, t1 = [], []
    for seed in range(seeds):
        net = m0.Network(m0.Config(), seed=seed)
        net.converged_cluster(n)
        x = net.create()
        assert net.join(x, 0) == 1
        t = net.first_tick(lambda: all(x in a.views for a in net.agents) and len(net.agents[x].views) == n + 1, 400)
        assert t is not None
        t0.append(t)
        p = OraclePool(lan_config(hostemu_lib, capacity=n + 1, n_initial=n, seed=3000 + seed, phase_group=1))
        y = p.member_add()
        assert p.join(y, [0]) == 1
        t = p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 400, 1)
        assert t != NEVER and len(p.members(y)) == n + 1
        t1.append(t)
    assert abs(st.mean(t0) - st.mean(t1)) <= 2.0, (t0, t1)


def test_stranded_broadcast_coverage_matches_full_fidelity_model(hostemu_lib):
    """50 % loss and RetransmitMult 1: gossip alone reaches only part of 400 agents — the same part,
    on average, in both models (the epidemic's final size is a property of the protocol)."""
    n, seeds = 400, 10
    c0, c1 = [], []
    for seed in range(seeds):
        net = m0.Network(m0.Config(loss=0.5, retransmit_mult=1), seed=seed)
        net.converged_cluster(n)
        key = net.user_event(7, b"e", b"p")
        net.step(120)
        c0.append(sum(1 for a in net.agents if any(k == key for _, k in a.delivered)) / n)
        p = OraclePool(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=4000 + seed, phase_group=1,
                                  packet_loss_ppm=500000, retransmit_mult=1))
        slot = p.user_event(7, b"e", b"p", False)
        p.step(120)
        c1.append(p.rumor_info(slot)["heard_count"] / n)
    assert 0.3 < st.mean(c0) < 1.0 and abs(st.mean(c0) - st.mean(c1)) <= 0.08, (c0, c1)


def test_piggyback_on_probe_traffic_shifts_dissemination_by_about_one_tick():
    """[U] memberlist/net.go sendMsg: pings, acks, indirect pings and nacks carry queued broadcasts and
    count as transmissions.  M1 and the CUDA path do not model it (probes are pull-evaluated: there is
    no probe packet to ride on); M0 can (Config.piggyback).  What the omission costs, measured here over
    8 seeds at 200 and 400 agents: the retransmit budget is conserved exactly (every agent still
    transmits an event RetransmitMult*ceil(log10(n+1)) times), 8-16 % of those transmissions move from
    gossip packets to probe packets, and the event reaches everybody 0.3-2 ticks (3-12 %) sooner — so
    M1's ticks-to-convergence are conservative by about one tick."""
    import statistics
    for n, limit in ((200, 12), (400, 12)):
        res = {}
        for piggy in (False, True):
            ticks, gossip, ridden = [], [], []
            for seed in range(1, 9):
                net = m0.Network(m0.Config(piggyback=piggy), seed=seed)
                net.converged_cluster(n)
                net.step(7)
                key = net.user_event(3, b"deploy", b"x" * 8)
                t0 = net.now
                t = net.first_tick(lambda: all(any(k == key for _, k in a.delivered) for a in net.up_agents()), 400)
                assert t is not None
                net.step(80)                                      # drain the queues
                ticks.append(t - t0)
                gossip.append(net.stats["msgs"])
                ridden.append(net.stats["piggyback_msgs"])
                assert net.stats["msgs"] + net.stats["piggyback_msgs"] == n * limit   # budget conserved
            res[piggy] = (statistics.mean(ticks), statistics.mean(gossip), statistics.mean(ridden))
        assert res[False][2] == 0
        share = res[True][2] / (n * limit)
        assert 0.06 < share < 0.2, share
        shift = res[False][0] - res[True][0]
        assert 0.0 < shift < 2.5, (n, res)


def test_dissemination_at_1600_agents_matches_the_projection(hostemu_lib):
    """The cross-check at the largest size the Python model runs in seconds (1 600 agents, 2.6 M view
    entries): both models transmit a user event exactly n * RetransmitMult * ceil(log10(n+1)) times and
    reach every agent within a tick or two of each other (3 seeds each; M0 and M1 draw differently, so
    the times are compared as means)."""
    n, limit = 1600, 16
    t_m0, t_m1 = [], []
    for seed in (1, 2, 3):
        net = m0.Network(m0.Config(), seed=seed)
        net.converged_cluster(n)
        net.step(5)
        key = net.user_event(3, b"deploy", b"x" * 8)
        s = net.now
        t = net.first_tick(lambda: all(any(k == key for _, k in a.delivered) for a in net.up_agents()), 300)
        assert t is not None
        net.step(80)
        assert net.stats["msgs"] == n * limit
        t_m0.append(t - s)
        p = Pool(lan_config(hostemu_lib, capacity=n, n_initial=n, seed=seed, phase_group=1), hostemu_lib)
        p.step(5)
        slot = p.user_event(3, b"deploy", b"x" * 8, False)
        s = p.now
        t = p.run_until(PRED_RUMOR_CONVERGED, slot, 300, 1)
        assert t != NEVER
        p.step(80)
        assert p.stats()["rumors_sent"] == n * limit and p.rumor_info(slot)["heard_count"] == n
        t_m1.append(t - s + 1)
    assert abs(st.mean(t_m0) - st.mean(t_m1)) <= 3.0, (t_m0, t_m1)
    assert 10 <= min(t_m0 + t_m1) and max(t_m0 + t_m1) <= 24


def _false_dead_pairs(net):
    up = net.up_agents()
    return sum(1 for a in up for b in up if a is not b and a.views[b.id].state == m0.DEAD)


def test_handle_reconnect_only_matters_where_running_members_stay_declared_dead():
    """serf's handleReconnect (M0 only, `Config.reconnect_interval`): with memberlist's defaults (TCP fallback
    ping) nobody running is ever declared dead at 30 % loss, so the ticker finds no failed member and does
    nothing — the regime of every BASELINE config, which is why M1 and the CUDA path leave it out.  Where
    running members do stay declared dead (no TCP fallback, 60 % loss: the accused never hears the rumor) the
    reconnect's push-pull is what heals the lists."""
    quiet = m0.Network(m0.Config(loss=0.30, reconnect_interval=300), seed=5)
    quiet.converged_cluster(60)
    quiet.step(1500)
    assert sum(a.stats["reconnect_attempts"] for a in quiet.agents) == 0 and _false_dead_pairs(quiet) == 0
    ends = {}
    for rc in (0, 300):
        tot = contacts = 0
        for seed in (1, 2, 3):
            net = m0.Network(m0.Config(loss=0.60, disable_tcp=True, reconnect_interval=rc), seed=seed)
            net.converged_cluster(60)
            net.step(2400)
            tot += _false_dead_pairs(net)
            contacts += sum(a.stats["reconnect_contacts"] for a in net.agents)
        ends[rc] = (tot, contacts)
    assert ends[0][1] == 0 and ends[300][1] > 0, ends
    assert ends[0][0] > 0 and ends[300][0] < ends[0][0] * 0.7, ends      # healed lists: clearly fewer stale Dead entries
