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


def check_invariants(p, prev=None, where=""):
    """Size-independent invariants of the pool state (no oracle needed: usable at BASELINE's full
    sizes).  `prev` = the dict returned by an earlier call on the same pool: monotone quantities
    (incarnations, Lamport clocks) must not have gone backwards.  Returns the new snapshot."""
    s = p.stats()
    n, now = s["n_members"], p.now
    if n == 0:
        return {"n": 0}
    act = active_mask(p)
    col = {name: p.column(name) for name in ("key", "meta", "due", "cursor", "probe_tgt", "ltime_member",
                                             "ltime_event", "heard", "queued", "tx", "sus_from")}
    key, meta = col["key"][:n], col["meta"][:n]
    truth, rank, inc = key & 3, (key >> 2) & 3, key >> 5
    exists, up = truth != 0, truth == 1
    stage, aw = (meta >> 3) & 3, meta & 7
    heard, queued = col["heard"][:n] & act, col["queued"][:n] & act
    limit = s["retransmit_limit"]

    def ok(cond, msg):
        assert bool(np.all(cond)), f"invariant violated {where}: {msg} (rows {np.nonzero(~np.asarray(cond))[0][:5].tolist()})"

    ok(inc[exists] >= 1, "an existing member has incarnation >= 1")
    ok(stage <= 2, "probe stage is IDLE / WAIT_T / WAIT_P")
    ok(aw[exists] < 8, "awareness below AwarenessMaxMultiplier")
    ok((queued[up] & ~heard[up]) == 0, "a member only re-broadcasts what it has heard")
    ok(col["due"][:n][up] >= now, "no running member has a probe action in the past")
    probing = up & (stage != 0)
    ok(col["probe_tgt"][:n][probing] < n, "an in-flight probe has a real target")
    ok(col["ltime_member"][:n][exists] >= 1, "member clock starts at 1")
    ok(col["ltime_event"][:n][exists] >= 1, "event clock starts at 1")
    for r in range(GSIM_MAX_RUMORS):
        if not (act >> r) & 1:
            continue
        tx = col["tx"][r][:n].astype(np.int64)
        has = ((heard >> r) & 1).astype(bool) & up
        q = ((queued >> r) & 1).astype(bool) & up
        ok(tx[has] <= limit, f"rumor {r}: transmits never exceed the retransmit limit")
        ok(tx[q] < limit, f"rumor {r}: a queued broadcast still has budget")
        info = p.rumor_info(r)
        assert info["heard_count"] == int(has.sum()), f"{where}: rumor {r} heard_count {info['heard_count']} != {int(has.sum())}"
        assert info["queued_count"] == int(q.sum()), f"{where}: rumor {r} queued_count"
    suspects = exists & (rank == 1)
    ok(col["sus_from"][0][:n][suspects] != 0xFFFFFFFF, "a suspect has a first accuser")
    assert s["n_up"] == int(up.sum()) and s["n_view_dead"] == int((exists & (rank == 2)).sum()), f"{where}: recount"
    snap = {"n": n, "inc": inc.copy(), "lm": col["ltime_member"][:n].copy(), "le": col["ltime_event"][:n].copy(),
            "exists": exists.copy()}
    if prev and prev.get("n"):
        m = prev["n"]
        both = prev["exists"] & exists[:m]
        ok(inc[:m][both] >= prev["inc"][both], "incarnations never decrease")
        ok(col["ltime_member"][:m][both] >= prev["lm"][both], "member clocks never decrease")
        ok(col["ltime_event"][:m][both] >= prev["le"][both], "event clocks never decrease")
    return snap
