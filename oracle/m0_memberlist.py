"""m0_memberlist.py — TEST INFRASTRUCTURE.  "M0" of SURVEY.md §7/§8d: a small, slow, FULL-FIDELITY
restatement of memberlist + serf's piggyback in which every agent keeps its OWN view of every other
agent (nodeState per observer, suspicion timers per (observer, subject), real broadcast queues
carrying real messages), i.e. the O(N^2) model that oracle/oracle.cpp ("M1") and the CUDA path
project onto one shared record per subject.  Pure Python, for N up to a few hundred.

It exists to check the PROJECTION, not to be compared bit for bit: M0 and M1 consume randomness
differently.  tests/test_m0_crosscheck.py compares what must agree — eventual outcomes, exact
counts where the protocol fixes them, and detection / dissemination times within the bounds and
distributions the protocol implies.

Restated from the published behaviour of ([U] = un-vendored upstream module, go.mod:80,85):
  [U] memberlist/state.go   probe, probeNode, gossip, aliveNode, suspectNode, deadNode, refute,
                            pushPull/mergeState (join only here)
  [U] memberlist/suspicion.go, awareness.go, queue.go (TransmitLimitedQueue), util.go
  [U] serf/serf.go          handleUserEvent, handleNodeJoinIntent/LeaveIntent, Leave; lamport.go
Only tests/ may import this file.  PARITY UNPINNED against the Go implementation.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field

ALIVE, SUSPECT, DEAD, LEFT = 0, 1, 2, 3


@dataclass
class Config:
    probe_interval: int = 10       # ticks (LAN: 1 s at tau = 100 ms)
    probe_timeout: int = 5
    gossip_interval: int = 2
    gossip_nodes: int = 3
    indirect_checks: int = 3
    retransmit_mult: int = 4
    suspicion_mult: int = 4
    suspicion_max_timeout_mult: int = 6
    awareness_max: int = 8
    gossip_to_the_dead: int = 300
    udp_budget: int = 1398
    loss: float = 0.0
    disable_tcp: bool = False
    event_buffer: int = 512
    tick_seconds: float = 0.1
    # [U] memberlist/net.go sendMsg / encodeAndSendMsg: every ping, ack, indirect ping, forwarded ack and
    # nack is sent as a compound packet with whatever broadcasts still fit (getBroadcasts), and carrying
    # them counts as a transmission.  Off by default: M1 and the CUDA path do not model it (probes are
    # pull-evaluated, there is no probe packet to ride on); tests/test_m0_crosscheck.py measures what
    # that omission costs.
    piggyback: bool = False
    ping_size: int = 40            # a ping / ack / nack with its node names, before any piggybacked broadcast
    # [U] serf/serf.go handleReconnect / reconnect: every ReconnectInterval (30 s) an agent with failed
    # members in its list draws one with probability failed / alive and tries memberlist.Join on it (a TCP
    # push-pull): a member that is really gone does not answer; one that was wrongly declared dead learns
    # of it from the merged state and refutes.  0 = off (M1 and the CUDA path do not model it;
    # tests/test_m0_crosscheck.py measures what it changes).
    reconnect_interval: int = 0


def retransmit_limit(mult, n):
    return mult * int(math.ceil(math.log10(n + 1.0)))


def suspicion_timeout_ticks(cfg, mult, n):
    scale = max(1.0, math.log10(max(1.0, float(n))))
    seconds = mult * int(scale * 1000.0) / 1000.0 * (cfg.probe_interval * cfg.tick_seconds)
    return seconds


def remaining_suspicion_seconds(confirmations, k, min_s, max_s):
    if k < 1:
        return min_s
    frac = math.log(confirmations + 1.0) / math.log(k + 1.0)
    raw = max_s - frac * (max_s - min_s)
    return max(min_s, math.floor(1000.0 * raw) / 1000.0)


@dataclass
class NodeState:
    inc: int = 0
    state: int = ALIVE
    change: int = 0


@dataclass
class Suspicion:
    start: int
    k: int
    min_s: float
    max_s: float
    confirmers: set = field(default_factory=set)


@dataclass
class Broadcast:
    kind: str          # alive / suspect / dead / join / leave / event
    node: int          # subject (alive/suspect/dead/intents) or origin (event)
    inc: int = 0
    frm: int = 0
    ltime: int = 0
    key: tuple = ()    # event identity
    size: int = 40
    transmits: int = 0
    seq: int = 0


class Agent:
    def __init__(self, net, ident):
        self.net, self.id = net, ident
        self.up, self.leaving = True, False
        self.inc = 1
        self.views = {ident: NodeState(1, ALIVE, 0)}
        self.ring, self.ring_pos = [], 0
        self.awareness = 0
        self.suspicions = {}
        self.queue = []                 # memberlist broadcasts
        self.intents = []               # serf intent queue
        self.events_q = []              # serf user-event queue
        self.clock_member, self.clock_event = 1, 1
        self.event_min = 0
        self.seen_events = set()
        self.delivered = []             # (tick, key) user events handed to the application
        self.probe_phase = net.rng.randrange(net.cfg.probe_interval)
        self.gossip_phase = net.rng.randrange(net.cfg.gossip_interval)
        self.next_probe = None
        self.probe = None               # in-flight probe: dict(target, start, stage, nacks_expected, nacks)
        self.reconnect_phase = net.rng.randrange(net.cfg.reconnect_interval) if net.cfg.reconnect_interval else 0
        self.stats = dict(refutes=0, probes=0, failed_probes=0, reconnect_attempts=0, reconnect_contacts=0)

    # ---- helpers -----------------------------------------------------------------------------
    def n(self):
        return len(self.views)

    def enqueue(self, b, q=None):
        q = self.queue if q is None else q
        if b.kind in ("alive", "suspect", "dead"):   # named broadcast: invalidates older ones about the node
            q[:] = [x for x in q if x.node != b.node]
        b.seq = self.net.next_seq()
        q.append(b)

    def k_random(self, k, accept):
        """[U] util.go kRandomNodes: up to 3n draws with rejection."""
        names = list(self.views)
        out = []
        for _ in range(3 * len(names)):
            if len(out) >= k:
                break
            c = self.net.rng.choice(names)
            if c == self.id or c in out or not accept(c, self.views[c]):
                continue
            out.append(c)
        return out

    # ---- memberlist state machine -------------------------------------------------------------------
    def alive_node(self, node, inc, t, bootstrap=False):
        st = self.views.get(node)
        if node == self.id and not bootstrap:
            if inc > self.inc or (st and st.state != ALIVE and inc >= self.inc):   # somebody thinks otherwise: refute handled by suspect/dead
                pass
            return
        if st is None:
            self.views[node] = NodeState(inc, ALIVE, t)
            pos = self.net.rng.randrange(len(self.ring) + 1)      # inserted at a random ring offset
            self.ring.insert(pos, node)
            self.enqueue(Broadcast("alive", node, inc, size=self.net.alive_size))
            return
        if inc <= st.inc:
            return
        st.inc, changed = inc, st.state != ALIVE
        st.state = ALIVE
        if changed:
            st.change = t
        self.suspicions.pop(node, None)
        self.enqueue(Broadcast("alive", node, inc, size=self.net.alive_size))

    def refute(self, accused_inc, t):
        self.inc = max(self.inc + 1, accused_inc + 1)
        self.views[self.id].inc = self.inc
        self.awareness = min(self.awareness + 1, self.net.cfg.awareness_max - 1)
        self.stats["refutes"] += 1
        self.enqueue(Broadcast("alive", self.id, self.inc, size=self.net.alive_size))

    def suspect_node(self, node, inc, frm, t):
        st = self.views.get(node)
        if st is None or inc < st.inc:
            return
        if node == self.id:
            if not self.leaving:
                self.refute(inc, t)
            return
        if st.state == SUSPECT:
            s = self.suspicions.get(node)
            if s and frm not in s.confirmers and frm != node and len(s.confirmers) - 1 < s.k:
                s.confirmers.add(frm)
                self.enqueue(Broadcast("suspect", node, inc, frm))      # re-gossip the confirmation
            return
        if st.state != ALIVE:
            return
        st.state, st.inc, st.change = SUSPECT, inc, t
        cfg = self.net.cfg
        k = cfg.suspicion_mult - 2
        if self.n() - 2 < k:
            k = 0
        mn = suspicion_timeout_ticks(cfg, cfg.suspicion_mult, self.n())
        self.suspicions[node] = Suspicion(t, max(0, k), mn, cfg.suspicion_max_timeout_mult * mn, {frm})
        self.enqueue(Broadcast("suspect", node, inc, frm))

    def dead_node(self, node, inc, frm, t):
        st = self.views.get(node)
        if st is None or inc < st.inc or st.state in (DEAD, LEFT):
            return
        if node == self.id and not self.leaving:
            self.refute(inc, t)
            return
        st.inc, st.change = inc, t
        st.state = LEFT if frm == node else DEAD
        self.suspicions.pop(node, None)
        self.enqueue(Broadcast("dead", node, inc, frm))

    def check_suspicions(self, t):
        cfg = self.net.cfg
        for node, s in list(self.suspicions.items()):
            total = remaining_suspicion_seconds(len(s.confirmers) - 1, s.k, s.min_s, s.max_s)
            if (t - s.start) * cfg.tick_seconds >= total - 1e-9:
                st = self.views[node]
                self.dead_node(node, st.inc, self.id, t)

    # ---- serf --------------------------------------------------------------------------------------
    def handle_event(self, b, t):
        self.clock_event = max(self.clock_event, b.ltime + 1)
        if b.ltime < self.event_min:
            return
        if self.clock_event > self.net.cfg.event_buffer and b.ltime < self.clock_event - self.net.cfg.event_buffer:
            return
        if b.key in self.seen_events:
            return
        self.seen_events.add(b.key)
        self.delivered.append((t, b.key))
        self.enqueue(Broadcast("event", b.node, ltime=b.ltime, key=b.key, size=b.size), self.events_q)

    def handle_intent(self, b, t):
        self.clock_member = max(self.clock_member, b.ltime + 1)
        ident = (b.kind, b.node, b.ltime)
        if ident in self.seen_events:
            return
        self.seen_events.add(ident)
        self.enqueue(Broadcast(b.kind, b.node, ltime=b.ltime, size=40), self.intents)

    # ---- receive -------------------------------------------------------------------------------------
    def receive(self, msgs, t):
        if not self.up:
            return
        for b in msgs:
            if b.kind == "alive":
                self.alive_node(b.node, b.inc, t)
            elif b.kind == "suspect":
                self.suspect_node(b.node, b.inc, b.frm, t)
            elif b.kind == "dead":
                self.dead_node(b.node, b.inc, b.frm, t)
            elif b.kind == "event":
                self.handle_event(b, t)
            else:
                self.handle_intent(b, t)

    # ---- tickers --------------------------------------------------------------------------------------
    def tick(self, t):
        if not self.up:
            return
        cfg = self.net.cfg
        self.check_suspicions(t)
        if self.next_probe is None:
            self.next_probe = t + (self.probe_phase - t) % cfg.probe_interval
        self.run_probe(t)
        if t % cfg.gossip_interval == self.gossip_phase:
            self.gossip(t)
        if cfg.reconnect_interval and t % cfg.reconnect_interval == self.reconnect_phase:
            self.reconnect(t)

    def reconnect(self, t):
        """[U] serf/serf.go reconnect(): one random failed member, throttled by failed / alive."""
        failed = [n for n, st in self.views.items() if n != self.id and st.state == DEAD]
        if not failed:
            return
        alive = sum(1 for st in self.views.values() if st.state in (ALIVE, SUSPECT)) or 1
        if self.net.rng.random() > len(failed) / alive:
            return                                        # "forgoing reconnect for random throttling"
        target = self.net.agents[failed[self.net.rng.randrange(len(failed))]]
        self.stats["reconnect_attempts"] += 1
        if not target.up:
            return                                        # nobody answers the TCP connect
        self.stats["reconnect_contacts"] += 1
        self.net.push_pull(self, target, t)

    def run_probe(self, t):
        cfg, net = self.net.cfg, self.net
        p = self.probe
        if p and p["stage"] == "timeout" and t == p["start"] + cfg.probe_timeout:
            j = p["target"]
            relays = self.k_random(cfg.indirect_checks, lambda c, st: c != j and st.state == ALIVE)
            ok, nacks = False, 0
            for r in relays:
                ra, ja = net.agents[r], net.agents[j]
                l_req = net.lost()
                self.piggyback(r, t, l_req)                       # indirect-ping request i -> relay
                if not ra.up or l_req:
                    continue
                l_ping = net.lost()
                ra.piggyback(j, t, l_ping)                        # relay's ping -> target
                l_ack = net.lost() if (ja.up and not l_ping) else True
                if ja.up and not l_ping:
                    ja.piggyback(r, t, l_ack)                     # target's ack -> relay
                acked = ja.up and not l_ping and not l_ack
                l_back = net.lost()
                ra.piggyback(self.id, t, l_back)                  # forwarded ack, or the nack, relay -> i
                if acked:
                    if not l_back:
                        ok = True
                elif not l_back:
                    nacks += 1
            if not cfg.disable_tcp and net.agents[j].up:
                ok = True
            if ok:
                self.awareness = max(0, self.awareness - 1)
                self.probe = None
                self.next_probe = p["start"] + cfg.probe_interval
            else:
                p.update(stage="deadline", missed=(len(relays) - nacks) if relays else 1,
                         deadline=p["start"] + cfg.probe_interval * (self.awareness + 1))
        p = self.probe
        if p and p["stage"] == "deadline" and t == p["deadline"]:
            self.awareness = min(self.awareness + p["missed"], cfg.awareness_max - 1)
            self.stats["failed_probes"] += 1
            self.probe = None
            self.next_probe = t
            st = self.views[p["target"]]
            self.suspect_node(p["target"], p["inc"], self.id, t)
        if self.probe is None and self.next_probe is not None and t >= self.next_probe:
            target = self.next_ring_target()
            if target is None:
                self.next_probe = t + cfg.probe_interval
                return
            self.stats["probes"] += 1
            st = self.views[target]
            ta = net.agents[target]
            l_ping = net.lost()
            self.piggyback(target, t, l_ping)                     # ping i -> target
            l_ack = net.lost() if (ta.up and not l_ping) else True
            if ta.up and not l_ping:
                ta.piggyback(self.id, t, l_ack)                   # ack target -> i
            if ta.up and not l_ping and not l_ack:
                self.awareness = max(0, self.awareness - 1)
                self.next_probe = t + cfg.probe_interval
            else:
                self.probe = dict(target=target, start=t, stage="timeout", inc=st.inc)
                self.next_probe = None

    def next_ring_target(self):
        checked = 0
        while checked < len(self.ring) + 1:
            if self.ring_pos >= len(self.ring):
                self.net.rng.shuffle(self.ring)                      # resetNodes + shuffle at the wrap
                self.ring_pos = 0
                checked += 1
                if not self.ring:
                    return None
                continue
            c = self.ring[self.ring_pos]
            self.ring_pos += 1
            st = self.views.get(c)
            if c == self.id or st is None or st.state in (DEAD, LEFT):
                checked += 1
                continue
            return c
        return None

    def take_broadcasts(self, budget):
        """[U] TransmitLimitedQueue.GetBroadcasts through serf's delegate: memberlist's queue first, then
        intents, then user events; fewest transmits first, then longest, then newest; every message taken
        counts one transmission and retires at the retransmit limit."""
        limit = retransmit_limit(self.net.cfg.retransmit_mult, self.n())
        packet, used = [], 0
        for q, overhead in ((self.queue, 2), (self.intents, 3), (self.events_q, 3)):
            for b in sorted(q, key=lambda b: (b.transmits, -b.size, -b.seq)):
                if used + overhead + b.size > budget:
                    continue
                packet.append(b)
                used += overhead + b.size
                b.transmits += 1
            q[:] = [b for b in q if b.transmits < limit]
        return [Broadcast(b.kind, b.node, b.inc, b.frm, b.ltime, b.key, b.size) for b in packet]

    def piggyback(self, dst, t, lost):
        """A probe-traffic packet from this agent to `dst` leaves now: it carries queued broadcasts
        (and they count as transmitted whether or not the packet arrives)."""
        if not self.net.cfg.piggyback or not self.up:
            return
        msgs = self.take_broadcasts(self.net.cfg.udp_budget - self.net.cfg.ping_size)
        if msgs:
            self.net.stats["piggyback_packets"] += 1
            self.net.stats["piggyback_msgs"] += len(msgs)
            if not lost:
                self.net.deliver(dst, msgs, t + 1)

    def gossip(self, t):
        cfg, net = self.net.cfg, self.net
        if not (self.queue or self.intents or self.events_q):
            return
        peers = self.k_random(cfg.gossip_nodes, lambda c, st: st.state in (ALIVE, SUSPECT) or
                              (st.state == DEAD and t - st.change <= cfg.gossip_to_the_dead))
        for peer in peers:
            packet = self.take_broadcasts(cfg.udp_budget)
            if not packet:
                break
            net.stats["packets"] += 1
            net.stats["msgs"] += len(packet)
            if not net.lost():
                net.deliver(peer, packet, t + 1)


class Network:
    def __init__(self, cfg: Config, seed=1, alive_size=64):
        self.cfg, self.rng = cfg, random.Random(seed)
        self.agents, self.now = [], 0
        self.inflight = {}
        self.seq = 0
        self.alive_size = alive_size
        self.stats = dict(packets=0, msgs=0, piggyback_packets=0, piggyback_msgs=0)

    def next_seq(self):
        self.seq += 1
        return self.seq

    def lost(self):
        return self.cfg.loss > 0 and self.rng.random() < self.cfg.loss

    def deliver(self, dst, msgs, at):
        self.inflight.setdefault(at, []).append((dst, msgs))

    # ---- the serf surface ---------------------------------------------------------------------------------
    def create(self):
        a = Agent(self, len(self.agents))
        self.agents.append(a)
        return a.id

    def converged_cluster(self, n):
        """n agents that already know each other (all Alive, incarnation 1)."""
        for _ in range(n):
            self.create()
        for a in self.agents:
            for b in self.agents:
                if a is not b:
                    a.views[b.id] = NodeState(1, ALIVE, 0)
            a.ring = [b.id for b in self.agents if b is not a]
            self.rng.shuffle(a.ring)

    def join(self, x, seed_id, ignore_old=True):
        """memberlist.Join = push-pull with the seed, then serf broadcasts a join intent."""
        a, s = self.agents[x], self.agents[seed_id]
        if not s.up or not a.up or a is s:
            return 0
        t = self.now
        self.push_pull(a, s, t)
        if ignore_old:
            a.event_min = max(a.event_min, s.clock_event)
        lt = a.clock_member
        a.clock_member += 1
        a.handle_intent(Broadcast("join", a.id, ltime=lt), t)
        return 1

    def push_pull(self, a, s, t):
        """[U] memberlist/state.go pushPullNode + mergeState, both directions, and serf's clocks"""
        for src, dst in ((s, a), (a, s)):
            for node, st in list(src.views.items()):
                if st.state == ALIVE:
                    dst.alive_node(node, st.inc, t)
                elif st.state in (SUSPECT, DEAD):
                    dst.alive_node(node, st.inc, t) if node not in dst.views else None
                    dst.suspect_node(node, st.inc, src.id, t)        # remote Dead is only a suspicion
            dst.clock_member = max(dst.clock_member, src.clock_member)
            dst.clock_event = max(dst.clock_event, src.clock_event)

    def user_event(self, x, name, payload=b""):
        a = self.agents[x]
        lt = a.clock_event
        a.clock_event += 1
        key = (lt, name, payload)
        size = 1 + 1 + 7 + (5 + 1 + len(name)) + (8 + 1 + len(payload)) + 4
        a.handle_event(Broadcast("event", x, ltime=lt, key=key, size=size), self.now)
        a.clock_event = max(a.clock_event, lt + 1)
        return key

    def crash(self, x):
        self.agents[x].up = False

    def leave(self, x):
        a = self.agents[x]
        lt = a.clock_member
        a.clock_member += 1
        a.leaving = True
        a.handle_intent(Broadcast("leave", x, ltime=lt), self.now)
        a.views[x].state = LEFT
        a.enqueue(Broadcast("dead", x, a.inc, x))

    def step(self, ticks=1):
        for _ in range(ticks):
            t = self.now
            for dst, msgs in self.inflight.pop(t, []):
                self.agents[dst].receive(msgs, t)
            order = list(range(len(self.agents)))
            for i in order:
                self.agents[i].tick(t)
            self.now = t + 1

    # ---- observation ------------------------------------------------------------------------------------
    def state_of(self, observer, subject):
        st = self.agents[observer].views.get(subject)
        return None if st is None else st.state

    def up_agents(self):
        return [a for a in self.agents if a.up]

    def all_see(self, subject, state):
        return all(a.views.get(subject) and a.views[subject].state == state for a in self.up_agents() if a.id != subject)

    def first_tick(self, pred, max_ticks):
        for _ in range(max_ticks):
            if pred():
                return self.now
            self.step(1)
        return None
