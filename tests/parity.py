"""Field-by-field comparison of two pools (CUDA / host-emulation / oracle)."""
from __future__ import annotations

import numpy as np

from consul_b200._lib import GSIM_MAX_RUMORS
from consul_b200.pool import GsimError

ACC_BIT = 0x80000000


def active_mask(pool) -> int:
    m = 0
    for r in range(GSIM_MAX_RUMORS):
        try:
            pool.rumor_info(r)
            m |= 1 << r
        except GsimError:
            pass
    return m


def compare_stats(a, b, where=""):
    sa, sb = a.stats(), b.stats()
    # active_rows is a scheduling diagnostic of the CUDA implementation, not simulation state
    diffs = {k: (sa[k], sb[k]) for k in sa if sa[k] != sb[k] and k != "active_rows"}
    assert not diffs, f"stats differ {where}: {diffs}"


def compare_columns(a, b, where=""):
    sa = a.stats()
    n = sa["n_members"]
    act = active_mask(a)
    assert act == active_mask(b), f"active rumor masks differ {where}"
    col = {name: (a.column(name), b.column(name)) for name in (
        "key", "meta", "due", "cursor", "pass", "probe_tgt", "probe_inc", "sus_start", "sus_from",
        "change_tick", "ltime_member", "ltime_event", "event_min", "heard", "queued", "tx", "inbox")}
    key = col["key"][0][:n]
    truth, rank = key & 3, (key >> 2) & 3
    meta = col["meta"][0][:n]
    stage = (meta >> 3) & 3
    up = truth == 1
    probing = up & (stage != 0)

    def eq(name, mask=None, transform=None):
        x, y = col[name]
        x, y = x[..., :n], y[..., :n]
        if transform is not None:
            x, y = transform(x), transform(y)
        if mask is not None:
            x, y = np.where(mask, x, 0), np.where(mask, y, 0)
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)[:5]
            raise AssertionError(f"column {name} differs {where} at {bad.tolist()}: "
                                 f"{[ (x[tuple(i)], y[tuple(i)]) for i in bad]}")

    eq("key")
    eq("meta")
    eq("due", up)
    eq("cursor")
    eq("pass")
    eq("probe_tgt", probing)
    eq("probe_inc", probing)
    eq("sus_start", rank == 1)
    eq("sus_from", (rank == 1)[None, :])
    eq("change_tick", rank >= 2)
    eq("ltime_member")
    eq("ltime_event")
    eq("event_min")
    eq("heard", transform=lambda v: v & act)
    eq("queued", transform=lambda v: v & act)
    eq("inbox", truth != 0, transform=lambda v: v & (act | ACC_BIT))
    heard = col["heard"][0][:n] & act
    bits = ((heard[None, :] >> np.arange(GSIM_MAX_RUMORS, dtype=np.uint32)[:, None]) & 1).astype(bool)
    eq("tx", bits)


def compare_pools(a, b, where="", columns=True):
    assert a.now == b.now, f"tick differs {where}: {a.now} vs {b.now}"
    compare_stats(a, b, where)
    if columns:
        compare_columns(a, b, where)
    ha, hb = a.state_hash(), b.state_hash()
    assert ha == hb, f"state hash differs {where}: {ha} vs {hb}"
