"""WAN federation driver — BASELINE config 5.

Two WAN gossip pools A and B (Consul: one serf WAN pool per federation,
/root/reference/agent/consul/server_serf.go:187-213, wanfed.go:36-40), each grouped into `n_dcs`
synthetic datacenters with an asymmetric datacenter-to-datacenter latency matrix, and
`bridges_per_dc` bridge members per datacenter that belong to both pools.  A bridge that first
delivers a user event in one pool re-fires it into the other pool one tick later (the ForwardRPC
of agent/consul/internal_endpoint.go:839 followed by the remote side's UserEvent).

The driver is host-side control flow only: it works on anything with the Pool interface
(consul_b200.Pool, consul_b200.ShardedPool, or the test oracle), steps both pools in lock step one
tick at a time while an event is crossing, reads the bridges' EventCh (gsim_poll_events) and
calls gsim_user_event / gsim_rumor_inject on the other pool.
"""
from __future__ import annotations

import numpy as np

EVENT_USER = 5
TILE = 128


def c5_latency_matrix(n_dcs: int = 64) -> np.ndarray:
    """L[a][b] = 1 + ((7a + 13b) mod 5) ticks, L[a][a] = 1 (asymmetric)."""
    a = np.arange(n_dcs, dtype=np.int64)[:, None]
    b = np.arange(n_dcs, dtype=np.int64)[None, :]
    m = 1 + (7 * a + 13 * b) % 5
    m[np.arange(n_dcs), np.arange(n_dcs)] = 1
    return m.astype(np.uint8)


def bridge_ids(n_dcs: int, bridges_per_dc: int, n_members: int):
    """The first `bridges_per_dc` members of every datacenter's first tile (same ids in both pools)."""
    ids = [dc * TILE + k for dc in range(n_dcs) for k in range(bridges_per_dc)]
    if ids and max(ids) >= n_members:
        raise ValueError("pool too small for one tile per datacenter")
    return ids


class WanFederation:
    def __init__(self, pool_a, pool_b, n_dcs: int = 64, bridges_per_dc: int = 5, n_members=None,
                 latency=None):
        self.pools = [pool_a, pool_b]
        n = n_members if n_members is not None else min(pool_a.stats()["n_members"], pool_b.stats()["n_members"])
        self.bridges = bridge_ids(n_dcs, bridges_per_dc, n)
        self._is_bridge = set(self.bridges)
        lat = c5_latency_matrix(n_dcs) if latency is None else latency
        # pools sharded over several GPUs: one process per GPU, all running this same driver
        self._rank = getattr(pool_a, "rank", 0)
        self._world = getattr(pool_a, "world", 1)
        for p in self.pools:
            p.latency_set(lat)
            for b in self.bridges:
                p.member_watch(b, True)
            if self._rank == 0:
                p.poll_events()                # start from an empty log
        # one logical event = one rumor slot per pool, keyed by (name, payload)
        self.slots = {}                        # key -> [slot in A or None, slot in B or None]
        self.key_of = [{}, {}]                 # per pool: slot -> key
        self.forwarded = 0                     # bridge re-fires that the other pool accepted
        self.forwarded_into = [0, 0]           # ... by receiving pool

    def _share(self, obj):
        if self._world <= 1:
            return obj
        import torch.distributed as dist
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def fire(self, pool_index: int, member: int, name: bytes, payload: bytes) -> int:
        slot = self.pools[pool_index].user_event(member, name, payload, False)
        key = (name, payload)
        self.slots.setdefault(key, [None, None])[pool_index] = slot
        self.key_of[pool_index][slot] = key
        return slot

    def step(self, ticks: int = 1):
        for _ in range(ticks):
            for p in self.pools:
                p.step(1)
            # deliveries logged at the tick that just ran are re-fired now, i.e. they enter the
            # other pool's tick t+1 — "one tick after first delivery"
            fired = []
            if self._rank == 0:            # a sharded pool's event log is served by rank 0
                for x, p in enumerate(self.pools):
                    for e in p.poll_events():
                        if e.type == EVENT_USER and e.observer in self._is_bridge and e.subject in self.key_of[x]:
                            fired.append((e.tick, x, e.observer, self.key_of[x][e.subject]))
            fired = self._share(fired)     # every rank issues the same calls (controller protocol)
            for _, x, member, key in sorted(fired):
                y = 1 - x
                slots = self.slots[key]
                if slots[y] is None:
                    self.fire(y, member, key[0], key[1])
                elif not self.pools[y].rumor_inject(slots[y], member):
                    continue
                self.forwarded += 1
                self.forwarded_into[y] += 1

    def converged(self, key) -> bool:
        slots = self.slots.get(key)
        if not slots or None in slots:
            return False
        for p, s in zip(self.pools, slots):
            info = p.rumor_info(s)
            if info["converged_tick"] == 0xFFFFFFFF:
                return False
        return True

    def run_until_converged(self, name: bytes, payload: bytes, max_ticks: int = 2000):
        """Ticks until every member of A and B has delivered the event, or None."""
        key = (name, payload)
        for _ in range(max_ticks):
            if self.converged(key):
                return max(p.rumor_info(s)["converged_tick"] for p, s in zip(self.pools, self.slots[key]))
            self.step(1)
        return None
