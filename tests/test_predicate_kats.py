"""Known-answer tests for the transition predicates and queue rules SURVEY.md §8c lists, checked on
the oracle AND on the kernel's row body (host emulation) separately — each must give the stated
answer on its own, not merely agree with the other.

  queue order     [U] memberlist/queue.go limitedBroadcast.Less: fewer transmits first, then the
                  longer message, then the newer one; GetBroadcasts fills one UDP packet per peer
  event window    [U] serf/serf.go handleUserEvent: Witness, then drop if LTime < clock - EventBuffer
  event de-dup    same (LTime, Name, Payload) is one event
  leave           [U] memberlist.Leave: dead{Node == From} => Left, never suspected
  refute          inc = max(inc + 1, accused + 1) — one bump per refutation
"""
import numpy as np
import pytest

from consul_b200.pool import Pool, lan_config
from oracle_binding import OraclePool


@pytest.fixture(params=["oracle", "kernel-body"])
def mk(request, hostemu_lib):
    if request.param == "oracle":
        return lambda cfg: OraclePool(cfg)
    return lambda cfg: Pool(cfg, hostemu_lib)


def sent_counts(p, member, slots):
    tx = p.column("tx")
    return [int(tx[s][member]) for s in slots]


def fire(p, member, payload_lens):
    return [p.user_event(member, b"ev%d" % k, b"p" * n, False) for k, n in enumerate(payload_lens)]


def gossip_once(p, member):
    """Advance to just past the member's next gossip tick (LAN: every 2 ticks)."""
    p.step(2)


def test_fewer_transmits_first(mk, hostemu_lib):
    """Budget for ONE message per packet, 3 peers: each of three queued messages goes out once
    rather than the first one three times."""
    cfg = lan_config(hostemu_lib, capacity=64, n_initial=64, seed=1, udp_buffer_size=2 + 3 + 60)
    p = mk(cfg)
    slots = fire(p, 9, [20, 20, 20])
    sizes = [p.rumor_info(s)["size_bytes"] for s in slots]
    assert max(sizes) + 3 <= 63 < 2 * (min(sizes) + 3)               # exactly one fits
    gossip_once(p, 9)
    assert sent_counts(p, 9, slots) == [1, 1, 1]


def test_longer_message_first(mk, hostemu_lib):
    """Equal transmits: the longest message is taken first.  2 peers, one message per packet:
    the shortest of three stays unsent."""
    cfg = lan_config(hostemu_lib, capacity=64, n_initial=64, seed=2, gossip_nodes=2, udp_buffer_size=2 + 3 + 80)
    p = mk(cfg)
    slots = fire(p, 9, [10, 40, 25])
    gossip_once(p, 9)
    assert sent_counts(p, 9, slots) == [0, 1, 1]


def test_newer_message_first_on_ties(mk, hostemu_lib):
    """Equal transmits and length: the newest goes first.  1 peer, one message per packet."""
    cfg = lan_config(hostemu_lib, capacity=64, n_initial=64, seed=3, gossip_nodes=1, udp_buffer_size=2 + 3 + 60)
    p = mk(cfg)
    slots = fire(p, 9, [20, 20, 20])
    gossip_once(p, 9)
    assert sent_counts(p, 9, slots) == [0, 0, 1]
    gossip_once(p, 9)
    assert sent_counts(p, 9, slots) == [0, 1, 1]


def test_packet_takes_everything_that_fits(mk, hostemu_lib):
    """GetBroadcasts keeps filling the packet: with room for two, each peer gets two messages."""
    cfg = lan_config(hostemu_lib, capacity=64, n_initial=64, seed=4, gossip_nodes=1, udp_buffer_size=2 + 2 * (3 + 60))
    p = mk(cfg)
    slots = fire(p, 9, [20, 20, 20])
    gossip_once(p, 9)
    assert sorted(sent_counts(p, 9, slots)) == [0, 1, 1]
    s = p.stats()
    assert s["gossip_packets"] == 1 and s["rumors_sent"] == 2


def test_memberlist_broadcasts_before_serf_events(mk, hostemu_lib):
    """[U] serf/delegate.go GetBroadcasts: memberlist's own broadcasts, then intents, then user
    events.  One message per packet, one peer: the joiner's alive message leaves before its join
    intent, and both before an older user event."""
    cfg = lan_config(hostemu_lib, capacity=70, n_initial=64, seed=5, gossip_nodes=1, udp_buffer_size=2 + 3 + 40)
    p = mk(cfg)
    ev = p.user_event(0, b"e", b"p" * 8, False)
    x = p.member_add(alive_msg_size=40)
    p.join(x, [0], False)                # no ignore_old: the joiner takes the seed's event too
    alive = [r for r in range(30) if _kind(p, r) == 1][0]
    intent = [r for r in range(30) if _kind(p, r) == 2][0]
    limit = p.stats()["retransmit_limit"]
    assert limit == 8                                                        # 4 * ceil(log10(65 + 1))
    heard_x = int(p.column("heard")[x])
    assert all((heard_x >> r) & 1 for r in (alive, intent, ev))              # the joiner queues all three
    seen = []
    for _ in range(3 * limit):
        gossip_once(p, x)
        seen.append(sent_counts(p, x, [alive, intent, ev]))
    # memberlist drains its own queue first (every packet), the delegate only gets what is left:
    # the alive message goes out `limit` times, then the intent `limit` times, then the event
    want = [[k, 0, 0] for k in range(1, limit + 1)] + [[limit, k, 0] for k in range(1, limit + 1)] + \
           [[limit, limit, k] for k in range(1, limit + 1)]
    assert seen == want


def _kind(p, r):
    try:
        return p.rumor_info(r)["kind"]
    except Exception:
        return 0


def test_event_window_and_dedup(mk, hostemu_lib):
    """EventBuffer = 4: after witnessing LTime 7 (clock 8) an event with LTime 3 is too old
    (3 < 8 - 4), LTime 4 is still accepted."""
    cfg = lan_config(hostemu_lib, capacity=16, n_initial=16, seed=6, event_buffer=4)
    p = mk(cfg)
    slots = [p.user_event(0, b"n%d" % k, b"", False) for k in range(7)]      # LTimes 1..7
    assert [p.rumor_info(s)["ltime"] for s in slots] == [1, 2, 3, 4, 5, 6, 7]
    assert int(p.column("ltime_event")[0]) == 8
    assert p.rumor_inject(slots[6], 5) is True                               # LTime 7
    assert int(p.column("ltime_event")[5]) == 8
    assert p.rumor_inject(slots[2], 5) is False                              # LTime 3: dropped
    assert p.rumor_inject(slots[3], 5) is True                               # LTime 4: kept
    heard5 = int(p.column("heard")[5])
    assert (heard5 >> slots[2]) & 1 == 0 and (heard5 >> slots[3]) & 1 == 1
    assert p.rumor_inject(slots[3], 5) is False                              # already delivered
    # same (LTime, Name, Payload) from another member with the same clock is the same event
    q = mk(cfg)
    a = q.user_event(1, b"x", b"y", False)
    b = q.user_event(2, b"x", b"y", False)
    c = q.user_event(3, b"x", b"z", False)
    assert a == b and c != a
    assert [int(v) for v in q.column("ltime_event")[1:4]] == [2, 2, 2]


def test_leave_is_left_never_suspected_and_refute_bumps_once(mk, hostemu_lib):
    cfg = lan_config(hostemu_lib, capacity=300, n_initial=300, seed=7, packet_loss_ppm=450000, disable_tcp_pings=1)
    p = mk(cfg)
    p.leave(4)
    key4 = int(p.column("key")[4])
    assert (key4 >> 2) & 3 == 3                                              # Left at once
    p.step(400)
    s = p.stats()
    assert s["refutes"] > 0
    inc = p.column("key")[:300] >> 5
    # every refutation bumps exactly one incarnation: sum(inc - 1) == refutes
    assert int((inc.astype(np.int64) - 1).sum()) == s["refutes"]
    key4 = int(p.column("key")[4])
    assert (key4 >> 2) & 3 == 3 and key4 >> 5 == 1                           # the leaver never refuted
