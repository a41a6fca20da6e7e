"""Known-answer tests that pin the oracle (and the product's host-side tables) to the upstream
formulas.  SURVEY.md §8c lists them; starred values there are recalled from upstream's own
test-suites ([U] memberlist/{util,suspicion}_test.go, serf/lamport_test.go), the rest follow
from the formula text pinned in /root/reference/agent/config/runtime.go:1310-1312,1328-1330.
"""
import ctypes as C

import pytest

from oracle_binding import oracle_lib

S, MS = 10**9, 10**6


def both():
    from consul_b200 import _lib
    o, g = oracle_lib(), _lib.lib()
    return [
        dict(name="oracle", rl=o.oracle_retransmit_limit, st=o.oracle_suspicion_timeout_ns,
             rs=o.oracle_remaining_suspicion_ns, pp=o.oracle_push_pull_scale_ns,
             lw=o.oracle_lamport_witness, ri=o.oracle_refute_incarnation, ph=o.oracle_philox4x32),
        dict(name="libgsim", rl=g.gsim_retransmit_limit, st=g.gsim_suspicion_timeout_ns,
             rs=g.gsim_remaining_suspicion_ns, pp=g.gsim_push_pull_scale_ns,
             lw=g.gsim_lamport_witness, ri=g.gsim_refute_incarnation, ph=g.gsim_philox4x32),
    ]


@pytest.mark.parametrize("impl", both(), ids=lambda d: d["name"])
def test_retransmit_limit(impl):
    rl = impl["rl"]
    assert rl(3, 0) == 0 and rl(3, 1) == 3 and rl(3, 99) == 6          # upstream util_test.go
    assert rl(4, 1_000_001) == 28 and rl(4, 999_998) == 24 and rl(4, 16_777_216) == 32
    assert rl(4, 4_000_000) == 28 and rl(4, 8_388_608) == 28          # BASELINE configs 3, 5
    for n, want in [(9, 4), (10, 8), (99, 8), (100, 12), (999_999, 24), (1_000_000, 28)]:
        assert rl(4, n) == want, n                                      # SURVEY §7 hard part 7


@pytest.mark.parametrize("impl", both(), ids=lambda d: d["name"])
def test_suspicion_timeout(impl):
    st = impl["st"]
    assert st(3, 10, S) == 3 * S and st(3, 100, S) == 6 * S and st(3, 1000, S) == 9 * S
    assert st(4, 3, 100 * MS) == 400 * MS                               # log10 floor at 1
    assert st(4, 4_000_000, S) == 26_408 * MS                           # BASELINE config 3
    assert st(2, 3, 100 * MS) == 200 * MS                               # server_test.go:221-237


@pytest.mark.parametrize("impl", both(), ids=lambda d: d["name"])
def test_remaining_suspicion_time(impl):
    rs = impl["rs"]                                                     # upstream suspicion_test.go
    assert rs(0, 3, 0, 2 * S, 30 * S) == 30 * S
    assert rs(1, 3, 2 * S, 2 * S, 30 * S) == 14 * S
    assert rs(2, 3, 3 * S, 2 * S, 30 * S) == 4810 * MS
    assert rs(3, 3, 4 * S, 2 * S, 30 * S) == -2 * S
    assert rs(0, 0, 0, 2 * S, 30 * S) == 2 * S                          # k < 1 -> min


@pytest.mark.parametrize("impl", both(), ids=lambda d: d["name"])
def test_push_pull_scale(impl):
    pp = impl["pp"]
    for n, mult in [(1, 1), (32, 1), (33, 2), (65, 3), (128, 3), (129, 4)]:
        assert pp(S, n) == mult * S, n


@pytest.mark.parametrize("impl", both(), ids=lambda d: d["name"])
def test_lamport_and_refute(impl):
    lw, ri = impl["lw"], impl["ri"]
    assert lw(0, 41) == 42 and lw(42, 41) == 42 and lw(42, 30) == 42 and lw(1, 1) == 2
    assert ri(5, 5) == 6 and ri(5, 9) == 10 and ri(5, 3) == 6


@pytest.mark.parametrize("impl", both(), ids=lambda d: d["name"])
def test_philox_random123_vectors(impl):
    def run(ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        impl["ph"](c, k, o)
        return [int(x) for x in o]
    # Random123 kat_vectors, philox4x32 10 rounds
    assert run([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_config_presets_match_consul_pins():
    """Defaults pinned by /root/reference/agent/config/runtime.go:1271-1413 and
    agent/consul/server_test.go:221-237, api/agent.go:299-303."""
    from consul_b200 import _lib
    from consul_b200.pool import lan_config, wan_config, consul_test_config
    L = _lib.lib()
    lan, wan, tst = lan_config(L), wan_config(L), consul_test_config(L)
    assert (lan.gossip_interval_ns, lan.gossip_nodes, lan.probe_interval_ns, lan.probe_timeout_ns,
            lan.retransmit_mult, lan.suspicion_mult) == (200 * MS, 3, S, 500 * MS, 4, 4)
    assert (wan.gossip_interval_ns, wan.gossip_nodes, wan.probe_interval_ns, wan.probe_timeout_ns,
            wan.retransmit_mult, wan.suspicion_mult) == (500 * MS, 3, 5 * S, 3 * S, 4, 6)
    assert (tst.probe_interval_ns, tst.probe_timeout_ns, tst.gossip_interval_ns,
            tst.suspicion_mult) == (100 * MS, 50 * MS, 100 * MS, 2)
    assert lan.leave_propagate_delay_ns == 3 * S                        # libserf/serf.go:33
    assert lan.reconnect_timeout_ns == 72 * 3600 * S                    # consul/config.go:622-623


def _classic_feistel(seed, n, member, pas, position):
    """The balanced 4-round swap-and-XOR network of round 1 (even domain widths), restated in Python."""
    M = 0xFFFFFFFF
    def fmix(x):
        x ^= x >> 16; x = (x * 0x85EBCA6B) & M; x ^= x >> 13; x = (x * 0xC2B2AE35) & M; x ^= x >> 16
        return x
    lo, hi = seed & M, seed >> 32
    k0 = fmix((member * 0x9E3779B1 + pas * 0x85EBCA77 + lo) & M)
    k1 = fmix(k0 ^ hi ^ 0xC2B2AE3D)
    keys = [k0, k1, (k0 * 0x9E3779B1 + k1) & M, ((k1 * 0x85EBCA77) & M) ^ k0]
    bits = 2
    while (1 << bits) < n:
        bits += 1
    assert bits % 2 == 0
    hb, mask, x = bits // 2, (1 << (bits // 2)) - 1, position
    while True:
        left, right = x >> hb, x & mask
        for k in keys:
            f = ((right + k) * 0x9E3779B1) & M
            f ^= f >> 15; f = (f * 0x85EBCA77) & M; f ^= f >> 13
            left, right = right, left ^ (f & mask)
        x = (left << hb) | right
        if x < n:
            return x


def test_probe_ring_is_a_permutation_for_every_width(hostemu_lib):
    """The probe ring (keyed Feistel with cycle walking) visits every entry exactly once per pass for member
    lists of any length — domain widths even and odd (an odd width uses halves that differ by one bit) —,
    product == oracle entry by entry, and for even widths it is still the classic balanced network."""
    from oracle_binding import oracle_lib
    O, L = oracle_lib(), hostemu_lib
    for n in (1, 2, 3, 4, 5, 9, 16, 17, 100, 1000, 2048, 2049, 4097, 5000, 8192, 8193, 20000):
        for member, pas in ((0, 0), (7, 3)):
            got = [L.gsim_ring_entry(0x5EED0001, n, member, pas, p) for p in range(n)]
            assert sorted(got) == list(range(n)), n
            assert got == [O.oracle_ring_entry(0x5EED0001, n, member, pas, p) for p in range(n)], n
            bits = max(2, (n - 1).bit_length())
            if bits % 2 == 0 and n <= 5000:
                assert got == [_classic_feistel(0x5EED0001, n, member, pas, p) for p in range(n)], n
            # the inverse (where an entry sits in the ring): quiet windows of a pristine pool rely on it
            assert [L.gsim_ring_position(0x5EED0001, n, member, pas, e) for e in got] == list(range(n)), n
    assert L.gsim_ring_entry(1, 10, 0, 0, 10) == 0xFFFFFFFF
    assert L.gsim_ring_position(1, 10, 0, 0, 10) == 0xFFFFFFFF


def test_fastmod_is_exact():
    """gs_fastmod (peer draws of the complete graph): ceil(2^64 / n) * x, high half times n == x % n for
    every 32-bit x and n — checked on the edges and on random pairs with Python integers."""
    import random
    rnd = random.Random(5)
    M = (1 << 64) - 1
    pairs = [(x, n) for n in (1, 2, 3, 7, 1000, 1000001, 4000000, 2**24, 2**31 - 1, 2**31, 2**32 - 1)
             for x in (0, 1, n - 1, n, n + 1, 2**32 - 1, 2**31)]
    pairs += [(rnd.getrandbits(32), rnd.randrange(1, 2**32)) for _ in range(20000)]
    for x, n in pairs:
        x &= 0xFFFFFFFF
        magic = (M // n + 1) & M
        low = (magic * x) & M
        assert (low * n) >> 64 == x % n, (x, n)
