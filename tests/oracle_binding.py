"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle/oracle.cpp).

`OraclePool` exposes the same methods as `consul_b200.pool.Pool` so a scenario can be
replayed on both and compared field by field.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from consul_b200._lib import (COLUMNS, GSIM_MAX_RUMORS, GSIM_MAX_SUSPICION_SLOTS, STAT_NAMES,
                              GsimConfig, GsimEvent, GsimMember, GsimMemberDesc, GsimRumorInfo,
                              GsimStats)
from consul_b200.pool import Event, GsimError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBORACLE = os.path.join(ROOT, "oracle", "liboracle.so")

_P, _u32, _u64, _i32, _sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t
_SIGS = [
    ("oracle_retransmit_limit", _u32, [_u32, _u32]),
    ("oracle_suspicion_timeout_ns", _u64, [_u32, _u32, _u64]),
    ("oracle_remaining_suspicion_ns", C.c_int64, [_u32, _u32, _u64, _u64, _u64]),
    ("oracle_push_pull_scale_ns", _u64, [_u64, _u32]),
    ("oracle_lamport_witness", _u32, [_u32, _u32]),
    ("oracle_refute_incarnation", _u32, [_u32, _u32]),
    ("oracle_ring_entry", _u32, [_u64, _u32, _u32, _u32, _u32]),
    ("oracle_philox4x32", None, [C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    ("oracle_create", _P, [C.POINTER(GsimConfig), _i32]),
    ("oracle_destroy", None, [_P]),
    ("oracle_threads", _i32, [_P]),
    ("oracle_set_threads", _i32, [_P, _i32]),
    ("oracle_config_default_lan", None, [C.POINTER(GsimConfig)]),
    ("oracle_config_default_wan", None, [C.POINTER(GsimConfig)]),
    ("oracle_config_consul_test", None, [C.POINTER(GsimConfig)]),
    ("oracle_member_add", _i32, [_P, C.POINTER(GsimMemberDesc), C.POINTER(_u32)]),
    ("oracle_join", _i32, [_P, _u32, C.POINTER(_u32), _sz, _i32, C.POINTER(_i32)]),
    ("oracle_leave", _i32, [_P, _u32]),
    ("oracle_crash_many", _i32, [_P, C.POINTER(_u32), _sz]),
    ("oracle_crash_fraction", _i32, [_P, _u32, _u32, C.POINTER(_u32)]),
    ("oracle_force_leave", _i32, [_P, _u32, _u32, _i32]),
    ("oracle_user_event", _i32, [_P, _u32, C.c_char_p, _sz, C.c_char_p, _sz, _i32, C.POINTER(_u32)]),
    ("oracle_rumor_inject", _i32, [_P, _u32, _u32, C.POINTER(_i32)]),
    ("oracle_latency_set", _i32, [_P, _u32, C.POINTER(C.c_uint8)]),
    ("oracle_graph_set", _i32, [_P, _u32, C.POINTER(_u32), C.POINTER(_u32)]),
    ("oracle_member_reconnect_timeout_set", _i32, [_P, _u32, C.c_uint64]),
    ("oracle_coordinate_get", _i32, [_P, _u32, C.POINTER(C.c_double)]),
    ("oracle_member_watch", _i32, [_P, _u32, _i32]),
    ("oracle_member_update", _i32, [_P, _u32, _u32, C.POINTER(_u32)]),
    ("oracle_step", _i32, [_P, _u32]),
    ("oracle_now", _u32, [_P]),
    ("oracle_run_until", _i32, [_P, _i32, _u32, _u32, _u32, C.POINTER(_u32)]),
    ("oracle_members", _i32, [_P, _u32, C.POINTER(GsimMember), _sz, C.POINTER(_sz)]),
    ("oracle_poll_events", _i32, [_P, C.POINTER(GsimEvent), _sz, C.POINTER(_sz)]),
    ("oracle_rumor_info_get", _i32, [_P, _u32, C.POINTER(GsimRumorInfo)]),
    ("oracle_rumor_retire", _i32, [_P, _u32]),
    ("oracle_stats_get", _i32, [_P, C.POINTER(GsimStats)]),
    ("oracle_state_hash", _i32, [_P, C.POINTER(_u64)]),
    ("oracle_column_read", _i32, [_P, _i32, _P, _sz, C.POINTER(_sz)]),
]

_LIB = None


def oracle_lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIBORACLE):
            raise OSError(f"{LIBORACLE} missing: run `python __graft_entry__.py`")
        lib = C.CDLL(LIBORACLE)
        for name, res, args in _SIGS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def oracle_config(preset: str = "lan", **kw) -> GsimConfig:
    """A preset config from the ORACLE's own restatement of the defaults (libgsim is not touched)."""
    c = GsimConfig()
    getattr(oracle_lib(), {"lan": "oracle_config_default_lan", "wan": "oracle_config_default_wan",
                           "consul_test": "oracle_config_consul_test"}[preset])(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


class OraclePool:
    def __init__(self, cfg: GsimConfig, threads: int = 1):
        self.lib = oracle_lib()
        self.cfg = cfg
        self.capacity = cfg.capacity
        self.h = self.lib.oracle_create(C.byref(cfg), threads)
        if not self.h:
            raise GsimError(-1, "oracle_create failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise GsimError(rc, "oracle")

    @property
    def threads(self):
        return self.lib.oracle_threads(self.h)

    def set_threads(self, n: int) -> int:
        return self.lib.oracle_set_threads(self.h, n)

    def member_add(self, alive_msg_size=0, watched=False, name_len=0, meta_len=0):
        d = GsimMemberDesc(alive_msg_size, 1 if watched else 0, name_len, meta_len)
        out = _u32()
        self._ck(self.lib.oracle_member_add(self.h, C.byref(d), C.byref(out)))
        return out.value

    def join(self, member, seeds, ignore_old=True):
        arr = (_u32 * len(seeds))(*seeds)
        n_ok = _i32()
        self._ck(self.lib.oracle_join(self.h, member, arr, len(seeds), int(ignore_old), C.byref(n_ok)))
        return n_ok.value

    def leave(self, member):
        self._ck(self.lib.oracle_leave(self.h, member))

    def crash(self, member):
        self.crash_many([member])

    def crash_many(self, ids):
        arr = (_u32 * len(ids))(*ids)
        self._ck(self.lib.oracle_crash_many(self.h, arr, len(ids)))

    def crash_fraction(self, ppm, salt=0):
        out = _u32()
        self._ck(self.lib.oracle_crash_fraction(self.h, ppm, salt, C.byref(out)))
        return out.value

    def force_leave(self, via, target, prune=False):
        self._ck(self.lib.oracle_force_leave(self.h, via, target, int(prune)))

    def user_event(self, member, name, payload, coalesce=False):
        out = _u32()
        self._ck(self.lib.oracle_user_event(self.h, member, name, len(name), payload, len(payload),
                                            int(coalesce), C.byref(out)))
        return out.value

    def rumor_inject(self, slot, member):
        out = C.c_int()
        self._ck(self.lib.oracle_rumor_inject(self.h, slot, member, C.byref(out)))
        return bool(out.value)

    def member_update(self, member, alive_msg_size=0):
        out = _u32()
        self._ck(self.lib.oracle_member_update(self.h, member, alive_msg_size, C.byref(out)))
        return out.value

    def graph_set(self, row_ptr, col_idx):
        import numpy as np
        if row_ptr is None:
            self._ck(self.lib.oracle_graph_set(self.h, 0, None, None))
            return
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        ci = np.ascontiguousarray(col_idx, dtype=np.uint32)
        self._ck(self.lib.oracle_graph_set(self.h, len(rp) - 1, rp.ctypes.data_as(C.POINTER(_u32)),
                                           ci.ctypes.data_as(C.POINTER(_u32))))

    def member_reconnect_timeout_set(self, member, timeout_ns):
        self._ck(self.lib.oracle_member_reconnect_timeout_set(self.h, member, timeout_ns))

    def coordinate(self, member):
        out = (C.c_double * 11)()
        self._ck(self.lib.oracle_coordinate_get(self.h, member, out))
        v = [float(x) for x in out]
        return v[:8], v[8], v[9], v[10]

    def member_watch(self, member, on=True):
        self._ck(self.lib.oracle_member_watch(self.h, member, int(on)))

    def latency_set(self, lat):
        import numpy as np
        if lat is None:
            self._ck(self.lib.oracle_latency_set(self.h, 0, None))
            return
        m = np.ascontiguousarray(lat, dtype=np.uint8)
        self._ck(self.lib.oracle_latency_set(self.h, m.shape[0], m.ctypes.data_as(C.POINTER(C.c_uint8))))

    def step(self, ticks=1):
        self._ck(self.lib.oracle_step(self.h, ticks))

    def run_until(self, predicate, arg=0, max_ticks=10000, check_every=16):
        out = _u32()
        self._ck(self.lib.oracle_run_until(self.h, predicate, arg, max_ticks, check_every, C.byref(out)))
        return out.value

    @property
    def now(self):
        return self.lib.oracle_now(self.h)

    def members(self, observer):
        n = _sz()
        self._ck(self.lib.oracle_members(self.h, observer, None, 0, C.byref(n)))
        buf = (GsimMember * max(1, n.value))()
        self._ck(self.lib.oracle_members(self.h, observer, buf, n.value, C.byref(n)))
        return [(m.id, m.status, m.incarnation, m.rank) for m in buf[: n.value]]

    def num_nodes(self, observer):
        return len(self.members(observer))

    def poll_events(self, cap=65536):
        buf = (GsimEvent * cap)()
        n = _sz()
        self._ck(self.lib.oracle_poll_events(self.h, buf, cap, C.byref(n)))
        return [Event(e.tick, e.type, e.subject, e.observer, e.ltime) for e in buf[: n.value]]

    def rumor_info(self, slot):
        out = GsimRumorInfo()
        self._ck(self.lib.oracle_rumor_info_get(self.h, slot, C.byref(out)))
        return {n: getattr(out, n) for n, _ in GsimRumorInfo._fields_}

    def rumor_retire(self, slot):
        self._ck(self.lib.oracle_rumor_retire(self.h, slot))

    def stats(self):
        s = GsimStats()
        self._ck(self.lib.oracle_stats_get(self.h, C.byref(s)))
        out = {n: int(s.counters[i]) for i, n in enumerate(STAT_NAMES)}
        for n, _ in GsimStats._fields_:
            if n == "counters":
                continue
            v = getattr(s, n)
            out[n] = list(v) if n == "suspicion_ticks" else int(v)
        return out

    def state_hash(self):
        out = (_u64 * 4)()
        self._ck(self.lib.oracle_state_hash(self.h, out))
        return tuple(int(x) for x in out)

    def column(self, name):
        cap = self.capacity
        if name == "tx":
            arr = np.zeros((GSIM_MAX_RUMORS, cap), dtype=np.uint8)
        elif name == "sus_from":
            arr = np.zeros((GSIM_MAX_SUSPICION_SLOTS, cap), dtype=np.uint32)
        else:
            arr = np.zeros(cap, dtype=np.uint32)
        n = _sz()
        self._ck(self.lib.oracle_column_read(self.h, COLUMNS[name], arr.ctypes.data_as(C.c_void_p),
                                             arr.nbytes, C.byref(n)))
        return arr
