"""Wire formats: libgsim's encoders (consul_b200/csrc/gs_wire.h through the C ABI)
against hand-assembled bytes from the msgpack specification, against an independent msgpack
implementation (msgpack-python: use_bin_type=False is the old "raw" spec that memberlist and serf's
zero codec.MsgpackHandle{} writes, use_bin_type=True the str8/bin spec of Consul's
msgpackHandleUserEvent), and — as sizes — against the oracle's arithmetic and against what the pool
charges a broadcast in a gossip packet (rumor_info.size)."""
import ctypes as C
import random
import struct

import msgpack
import pytest

from consul_b200.pool import Pool, lan_config
from oracle_binding import OraclePool


def enc(lib, fn, *args, cap=4096):
    buf = C.create_string_buffer(cap)
    n = getattr(lib, fn)(buf, cap, *args)
    assert n <= cap
    assert getattr(lib, fn)(None, 0, *args) == n              # sizing call: same answer, nothing written
    return buf.raw[:n]


def old(obj):                                                 # memberlist / serf handle
    return msgpack.packb(obj, use_bin_type=False)


VSN = (C.c_uint8 * 6)(1, 5, 2, 2, 5, 4)


def test_known_answers(hostemu_lib):
    L = hostemu_lib
    # suspect{Incarnation: 3, Node: "b", From: "a"}: type 3, fixmap 3, fixraw keys in struct order
    assert enc(L, "gsim_wire_suspect", 3, b"b", b"a") == \
        bytes([3, 0x83, 0xAB]) + b"Incarnation" + bytes([3, 0xA4]) + b"Node" + bytes([0xA1]) + b"b" + \
        bytes([0xA4]) + b"From" + bytes([0xA1]) + b"a"
    assert enc(L, "gsim_wire_dead", 300, b"b", b"b")[:16] == bytes([5, 0x83, 0xAB]) + b"Incarnation" + bytes([0xCD, 0x01])
    # serf user event: serf type 3, map of 4
    assert enc(L, "gsim_wire_user_event", 1, b"deploy", 6, b"v1", 2, 0) == \
        bytes([3, 0x84, 0xA5]) + b"LTime" + bytes([1, 0xA4]) + b"Name" + bytes([0xA6]) + b"deploy" + \
        bytes([0xA7]) + b"Payload" + bytes([0xA2]) + b"v1" + bytes([0xA2]) + b"CC" + bytes([0xC2])
    # join / leave intents: serf types 1 / 0
    assert enc(L, "gsim_wire_join_intent", 70000, b"n1") == \
        bytes([1, 0x82, 0xA5]) + b"LTime" + bytes([0xCE, 0, 1, 0x11, 0x70, 0xA4]) + b"Node" + bytes([0xA2]) + b"n1"
    assert enc(L, "gsim_wire_leave_intent", 2, b"n1", 1)[-7:] == bytes([0xA5]) + b"Prune" + bytes([0xC3])
    # a 32-byte node name is raw16 (0xDA 0x00 0x20): there is no str8 without WriteExt
    name = b"x" * 32
    a = enc(L, "gsim_wire_alive", 1, name, bytes([10, 0, 0, 1]), 4, 8301, b"", 0, VSN)
    assert a[:2] == bytes([4, 0x86]) and bytes([0xA4]) + b"Node" + bytes([0xDA, 0x00, 0x20]) + name in a
    assert bytes([0xA4]) + b"Port" + bytes([0xCD, 0x20, 0x6D]) in a                       # 8301, uint16
    assert a.endswith(bytes([0xA3]) + b"Vsn" + bytes([0xA6, 1, 5, 2, 2, 5, 4]))
    # memberlist compound packet: type 7, count, big-endian u16 lengths, bodies
    m1, m2 = b"abc", b"defgh"
    ptrs = (C.c_void_p * 2)(C.cast(C.c_char_p(m1), C.c_void_p), C.cast(C.c_char_p(m2), C.c_void_p))
    lens = (C.c_size_t * 2)(3, 5)
    assert enc(L, "gsim_wire_compound", ptrs, lens, 2) == bytes([7, 2, 0, 3, 0, 5]) + m1 + m2
    # wanfed: big-endian u32 length, then the packet (agent/consul/wanfed/wanfed.go:112-121)
    assert enc(L, "gsim_wire_wanfed_frame", b"hello", 5) == struct.pack(">I", 5) + b"hello"
    # Consul's UserEvent payload (agent/user_event.go:27-52): tagged keys, omitempty, bin payload, str8 names
    assert enc(L, "gsim_wire_consul_user_event", b"id", b"deploy", b"xyz", 3, None, None, None, 1) == \
        bytes([0x84, 0xA2]) + b"ID" + bytes([0xA2]) + b"id" + bytes([0xA1]) + b"n" + bytes([0xA6]) + b"deploy" + \
        bytes([0xA1]) + b"p" + bytes([0xC4, 3]) + b"xyz" + bytes([0xA1]) + b"v" + bytes([1])
    long_name = b"n" * 40
    e = enc(L, "gsim_wire_consul_user_event", b"i", long_name, None, 0, b"^web", None, None, 1)
    assert e[0] == 0x84 and bytes([0xA1]) + b"n" + bytes([0xD9, 40]) + long_name in e and bytes([0xA2]) + b"nf" in e


def test_against_an_independent_msgpack(hostemu_lib):
    L = hostemu_lib
    rnd = random.Random(7)
    for _ in range(300):
        inc = rnd.choice([1, 5, 127, 128, 255, 256, 65535, 65536, 2**27])
        node = bytes(rnd.choice(b"abcdefgh.-") for _ in range(rnd.choice([1, 5, 31, 32, 33, 255, 256, 300])))
        frm = bytes(rnd.choice(b"xyz") for _ in range(rnd.choice([1, 31, 32])))
        meta = bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 31, 32, 200, 512])))
        lt = rnd.choice([0, 1, 127, 128, 70000, 2**32 + 5])
        addr = bytes([10, rnd.randrange(256), 0, 1])
        got = enc(L, "gsim_wire_alive", inc, node, addr, 4, 8301, meta, len(meta), VSN)
        assert got == bytes([4]) + old({b"Incarnation": inc, b"Node": node, b"Addr": addr, b"Port": 8301, b"Meta": meta,
                                        b"Vsn": bytes(VSN)})
        assert enc(L, "gsim_wire_suspect", inc, node, frm) == bytes([3]) + old({b"Incarnation": inc, b"Node": node, b"From": frm})
        assert enc(L, "gsim_wire_join_intent", lt, node) == bytes([1]) + old({b"LTime": lt, b"Node": node})
        assert enc(L, "gsim_wire_leave_intent", lt, node, 0) == bytes([0]) + old({b"LTime": lt, b"Node": node, b"Prune": False})
        assert enc(L, "gsim_wire_user_event", lt, node, len(node), meta, len(meta), 1) == \
            bytes([3]) + old({b"LTime": lt, b"Name": node, b"Payload": meta, b"CC": True})
        want = {"ID": "u-1", "n": node.decode()}
        if meta:
            want["p"] = meta
        want["v"] = 1
        assert enc(L, "gsim_wire_consul_user_event", b"u-1", node, meta, len(meta), None, None, None, 1) == \
            msgpack.packb(want, use_bin_type=True)


@pytest.mark.parametrize("name_len", [0, 4, 31, 32, 200])
def test_pool_charges_the_encoded_sizes(hostemu_lib, name_len):
    """What a broadcast costs in a gossip packet is the encoder's size (product) = the oracle's arithmetic."""
    L = hostemu_lib
    cfg = lan_config(L, capacity=40, n_initial=30, seed=9)
    pools = [Pool(cfg, L), OraclePool(cfg)]
    sizes = []
    for p in pools:
        x = p.member_add(name_len=name_len, meta_len=77)
        p.join(x, [0])
        ev = p.user_event(3, b"n" * 33, b"p" * 40, False)
        p.leave(5)
        upd = p.member_update(x)
        sizes.append({(p.rumor_info(r)["kind"], p.rumor_info(r)["subject"]): p.rumor_info(r)["size_bytes"]
                      for r in range(30) if _active(p, r)})
    assert sizes[0] == sizes[1]
    name = b"m" * name_len if name_len else b"node-30"
    assert sizes[0][(1, 30)] == len(enc(L, "gsim_wire_alive", 1, name, bytes(4), 4, 8301, b"t" * 77, 77, VSN))
    assert sizes[0][(2, 30)] == len(enc(L, "gsim_wire_join_intent", 1, name))
    assert sizes[0][(3, 5)] == len(enc(L, "gsim_wire_leave_intent", 1, b"node-5", 0))
    assert sizes[0][(4, 3)] == len(enc(L, "gsim_wire_user_event", 1, b"n" * 33, 33, b"p" * 40, 40, 0))
    assert sizes[0][(5, 30)] == len(enc(L, "gsim_wire_alive", 2, name, bytes(4), 4, 8301, b"", 0, VSN))


def _active(p, r):
    try:
        p.rumor_info(r)
        return True
    except Exception:
        return False
