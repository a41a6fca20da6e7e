"""Quiet windows: up to ProbeInterval ticks of a quiet pool in one launch.  The
schedule must be invisible in the results — same digest, counters and columns as one launch per tick
and as the oracle — whatever happens around and inside the windows: joins, events, crashes (probe
failures are what bounds a window: the horizon), packet loss, leaves, steps of any length, restores.
The host emulation runs each row through ALL its ticks of a window before it looks at the next row,
in forward / reverse / odd-even row order: inside a window rows really are independent."""
import os
import subprocess
import sys

import pytest

from consul_b200.pool import (FLAG_NO_WINDOWS, PRED_ALL_RUMORS_CONVERGED, PRED_CRASHED_ALL_DEAD, Pool,
                              consul_test_config, lan_config, wan_config)
from oracle_binding import OraclePool
from parity import compare_pools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def trio(lib, cfg_fn, **kw):
    """the same pool with windows, without, and on the oracle"""
    flags = kw.pop("flags", 0)
    return [Pool(cfg_fn(lib, flags=flags, **kw), lib), Pool(cfg_fn(lib, flags=flags | FLAG_NO_WINDOWS, **kw), lib),
            OraclePool(cfg_fn(lib, flags=flags, **kw))]


def all3(pools, fn):
    out = [fn(p) for p in pools]
    assert out[0] == out[1] == out[2], out
    return out[0]


def check(pools, where):
    compare_pools(pools[0], pools[2], where + " (windows vs oracle)")
    compare_pools(pools[0], pools[1], where + " (windows vs single ticks)")


def test_steady_state_runs_in_windows(hostemu_lib):
    pools = trio(hostemu_lib, lan_config, capacity=5000, n_initial=5000, seed=21)
    for p in pools:
        p.step(500)
    check(pools, "steady 500")
    sc = pools[0].sched_counts()
    assert sc["window_ticks"] > 450 and sc["window_launches"] <= 60, sc     # 10 ticks per launch
    assert pools[1].sched_counts()["window_ticks"] == 0
    for k in (1, 3, 9, 10, 11, 27):                                         # windows of every length, partial tails
        for p in pools:
            p.step(k)
        check(pools, f"steady +{k}")


@pytest.mark.parametrize("n,seed", [(257, 31), (1000, 32)])
def test_pristine_pool_closed_form(hostemu_lib, n, seed, chunks=None):
    """Everybody up, alive, established, no loss: a launch covers 256 ProbeIntervals and a member's probes
    are advanced in closed form (gs_pristine_probes) — through the end of ring passes (n probes = 10 n ticks)
    and past each member's own ring entry, which the generic step has to skip.  Same digest, counters and
    columns as one launch per tick and as the oracle at every checkpoint."""
    pools = trio(hostemu_lib, lan_config, capacity=n + 1, n_initial=n, seed=seed)
    total = 0
    for chunk in chunks or (100, 2560, 7, 5000, 1, 2559, 12 * n):
        for p in pools:
            p.step(chunk)
        total += chunk
        check(pools, f"pristine after {total}")
    sc = pools[0].sched_counts()
    assert sc["window_ticks"] > total - 200 and sc["window_launches"] <= 4 + total // 2560 + 7, sc
    st = pools[0].stats()
    assert st["probes"] == st["acks"] and st["probes"] >= (total // 10 - 2) * n      # every probe a prompt ack
    # a member that is not established ends it: the joiner's alive rumor has to be heard first
    x = all3(pools, lambda p: p.member_add())
    assert all3(pools, lambda p: p.join(x, [0])) == 1
    cf0 = pools[0].sched_counts()["closed_form_ticks"]
    assert cf0 > total - 200
    for p in pools:
        p.step(3000)                      # ONE step: the alive rumor is retired (the joiner established) only at its end
    check(pools, "joined + 3000")
    # ... but the closed form still runs once the rumor has been heard by everybody: it stops in front of the
    # joiner's ring entry (whether a prober knows a pending member is the generic step's business)
    assert pools[0].sched_counts()["closed_form_ticks"] - cf0 > 2000, pools[0].sched_counts()
    for p in pools:
        p.step(2000)
    check(pools, "joined + 5000")


def test_closed_form_stops_at_a_member_somebody_has_not_heard_of(hostemu_lib):
    """A joiner whose alive rumor dies out before everybody has heard it (one peer per gossip tick, three
    transmissions) stays pending: members that have heard of it probe it, the others skip its ring entry.
    The closed form must leave that entry to the generic step."""
    pools = trio(hostemu_lib, lan_config, capacity=301, n_initial=300, seed=35, gossip_nodes=1, retransmit_mult=1)
    x = all3(pools, lambda p: p.member_add())
    assert all3(pools, lambda p: p.join(x, [0])) == 1
    for chunk in (200, 3100, 2900, 1):                       # every member passes the joiner's entry at least once
        for p in pools:
            p.step(chunk)
        check(pools, f"half-heard joiner +{chunk}")
    slot = []
    for r in range(30):
        try:
            if pools[0].rumor_info(r)["kind"] == 1:          # GSIM_RUMOR_ALIVE
                slot.append(r)
        except Exception:
            pass                                              # free slot
    assert len(slot) == 1
    heard = all3(pools, lambda p: p.rumor_info(slot[0])["heard_count"])
    assert 100 < heard < 301, heard                          # some have heard it, some never will
    sc = pools[0].sched_counts()
    assert sc["closed_form_ticks"] > 5000, sc


def test_join_cascade_then_windows(hostemu_lib):
    pools = trio(hostemu_lib, lan_config, capacity=3001, n_initial=3000, seed=22)
    x = all3(pools, lambda p: p.member_add())
    assert all3(pools, lambda p: p.join(x, [0])) == 1
    for p in pools:
        p.step(400)
    check(pools, "cascade + steady")
    sc = pools[0].sched_counts()
    assert sc["tick_launches"] >= 20 and sc["window_ticks"] >= 200, sc      # single ticks while the rumor runs, windows after
    y = all3(pools, lambda p: p.user_event(7, b"deploy", b"now", False))    # a host call ends the quiet period
    for p in pools:
        p.step(300)
    check(pools, "event + steady")
    assert all3(pools, lambda p: p.rumor_info(y)["heard_count"]) == 3001


@pytest.mark.parametrize("cfg_fn,ppm,ticks", [(lan_config, 20000, 700), (consul_test_config, 100000, 200), (wan_config, 30000, 900)])
def test_crashes_bound_the_windows(hostemu_lib, cfg_fn, ppm, ticks):
    """Crashed members make probes fail inside windows: the window that sees the failure lowers the
    horizon, the chain stops there, accusations and suspicion run as single ticks, and once every
    crashed member is Dead and the rumors have drained the pool is quiet again."""
    pools = trio(hostemu_lib, cfg_fn, capacity=4000, n_initial=4000, seed=23)
    for p in pools:
        p.step(40)
    crashed = all3(pools, lambda p: p.crash_fraction(ppm, 1))
    assert crashed > 0
    for chunk in (7, 64, 200, ticks):
        for p in pools:
            p.step(chunk)
        check(pools, f"crash wave +{chunk}")
    t = all3(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 4000, 16))
    for p in pools:
        p.step(300)
    check(pools, "after the wave")
    sc = pools[0].sched_counts()
    assert sc["window_ticks"] > 100 and sc["tick_launches"] > 30 and sc["horizon_scans"] >= 2, sc
    assert pools[0].stats()["probe_failures"] > 0


def test_lossy_pool_and_leave(hostemu_lib):
    """Random probe failures (5 % loss) keep lowering the horizon; results still equal."""
    pools = trio(hostemu_lib, lan_config, capacity=1200, n_initial=1200, seed=24, packet_loss_ppm=50000)
    for p in pools:
        p.step(333)
    check(pools, "lossy steady")
    all3(pools, lambda p: p.leave(17))
    for p in pools:
        p.step(444)
    check(pools, "lossy + leave")
    assert pools[0].stats()["nacks"] > 0


def test_snapshot_restore_inside_a_quiet_period(hostemu_lib):
    pools = trio(hostemu_lib, lan_config, capacity=2000, n_initial=2000, seed=25)
    for p in pools:
        p.step(123)
    blobs = [p.snapshot() for p in pools[:2]]
    for p in pools[:2]:
        p.step(57)
    for p, b in zip(pools[:2], blobs):
        p.restore(b)
    for p in pools[:2]:
        p.step(200)
    pools[2].step(200)
    check(pools, "restore + 200")


def test_window_schedule_is_independent_of_row_order():
    """reverse and odd/even row orders inside windows (each row runs all its ticks of a window first)"""
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from consul_b200 import _lib\n"
        "from consul_b200.pool import Pool, lan_config\n"
        "L = _lib.load(%r)\n"
        "p = Pool(lan_config(L, capacity=3001, n_initial=3000, seed=26), L)\n"
        "p.step(50); p.crash_fraction(30000, 2); p.step(40); x = p.member_add(); p.join(x, [1]); p.step(1500)\n"
        "s = p.stats(); s.pop('active_rows')\n"
        "print(p.state_hash(), sorted(s.items()), p.sched_counts()['window_ticks'] > 100)\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostemu", "libgsim_hostemu.so"))
    outs = []
    for order in ("0", "1", "2"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GSIM_HOSTEMU_ORDER=order),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] == outs[2] and outs[0].endswith("True"), outs


def test_long_launches_only_while_no_probe_can_fail(hostemu_lib):
    """On a quiet pool where every listed member runs (no loss, no slow link) no probe can go unanswered,
    so one launch covers many ProbeIntervals; one crashed member that is still listed alive ends that
    (launches of one ProbeInterval, bounded by the horizon, then single ticks), and once it is Dead —
    skipped by every probe ring — the long launches are back.  Same results as single ticks throughout."""
    pools = trio(hostemu_lib, lan_config, capacity=3000, n_initial=3000, seed=27)
    for p in pools:
        p.step(700)
    check(pools, "healthy 700")
    sc = pools[0].sched_counts()
    assert sc["window_ticks"] >= 650 and sc["window_launches"] <= 8, sc              # 320 ticks per launch
    all3(pools, lambda p: p.crash(1234))
    before = pools[0].sched_counts()
    for p in pools:
        p.step(60)
    check(pools, "one crashed, still listed")
    mid = pools[0].sched_counts()
    assert mid["tick_launches"] > before["tick_launches"]                           # the failed probe ended the windows
    t = all3(pools, lambda p: p.run_until(PRED_CRASHED_ALL_DEAD, 0, 3000, 32))
    for p in pools:
        p.step(800)
    check(pools, "dead, skipped by the rings")
    after = pools[0].sched_counts()
    assert after["window_ticks"] - mid["window_ticks"] >= 600
    assert after["window_launches"] - mid["window_launches"] <= 40, (mid, after)    # long launches again
