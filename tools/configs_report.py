"""BASELINE.json's second metric — ticks-to-full-convergence — for every config that fits one GPU,
one JSON line each (run on a B200; ~1 minute):

  C2  1 000 000 members + 1 joiner, seeds 0x5EED0001..3      ticks until every member lists the joiner
  C3  4 000 000 members, 10 % crashed at tick 0              first Dead, all crashed Dead, false positives
  C4  16 777 216 members, one user event                     ticks until every member delivered it
  C5  2 x 8 388 608 members, 64 DCs, C5 matrix, 5 bridges/DC ticks until both WAN pools delivered it

Usage: python tools/configs_report.py [c2 c3 c4 c5]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consul_b200 import (NEVER, PRED_ALL_RUMORS_CONVERGED, PRED_CRASHED_ALL_DEAD, PRED_RUMOR_CONVERGED, Pool,  # noqa: E402
                         WanFederation, lan_config, wan_config)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c2():
    for seed in (0x5EED0001, 0x5EED0002, 0x5EED0003):
        n = 1_000_000
        p = Pool(lan_config(capacity=n + 1, n_initial=n, seed=seed))
        x = p.member_add()
        p.join(x, [0])
        t0 = time.perf_counter()
        t = p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 600, 1)
        emit(config="C2", members=n + 1, seed=hex(seed), tau_ms=100, ticks_to_convergence=None if t == NEVER else t,
             wall_s=round(time.perf_counter() - t0, 3), digest="%016x" % p.state_hash()[0])
        p.close()


def c3():
    n = 4_000_000
    p = Pool(lan_config(capacity=n, n_initial=n, seed=0x5EED0002))
    crashed = p.crash_fraction(100000, 3)
    t0 = time.perf_counter()
    first = None
    while first is None and p.now < 400:
        p.step(1)
        if p.stats()["deads"]:
            first = p.now
    t = p.run_until(PRED_CRASHED_ALL_DEAD, 0, 4000, 16)
    s = p.stats()
    emit(config="C3", members=n, crashed=crashed, tau_ms=100, first_dead_tick=first,
         all_crashed_dead_tick=None if t == NEVER else t, false_positives=s["deads"] - crashed, refutes=s["refutes"],
         suspicion_ticks=s["suspicion_ticks"][:3], wall_s=round(time.perf_counter() - t0, 3))
    p.close()


def c4():
    n = 16 * 1024 * 1024
    p = Pool(lan_config(capacity=n, n_initial=n, seed=0x5EED0003))
    slot = p.user_event(0, b"deploy", b"x" * 32, False)
    t0 = time.perf_counter()
    t = p.run_until(PRED_RUMOR_CONVERGED, slot, 600, 1)
    lt = p.column("ltime_event")[:n]
    emit(config="C4", members=n, gpus=1, tau_ms=100, ticks_to_convergence=None if t == NEVER else t,
         min_event_clock=int(lt.min()), rumors_accepted=p.stats()["rumors_accepted"],
         wall_s=round(time.perf_counter() - t0, 3), digest="%016x" % p.state_hash()[0])
    p.close()


def c5():
    n = 8 * 1024 * 1024
    mk = lambda seed: Pool(wan_config(capacity=n, n_initial=n, seed=seed, mailbox_depth=8))  # noqa: E731
    fed = WanFederation(mk(0x5EED0051), mk(0x5EED0052), n_dcs=64, bridges_per_dc=5, n_members=n)
    fed.fire(0, 7, b"deploy", b"x" * 32)
    t0 = time.perf_counter()
    t = fed.run_until_converged(b"deploy", b"x" * 32, 600)
    emit(config="C5", members_per_pool=n, pools=2, dcs=64, bridges_per_dc=5, tau_ms=500, ticks_to_convergence=t,
         bridge_refires=fed.forwarded, refires_into=fed.forwarded_into,
         suspects=[p.stats()["suspects"] for p in fed.pools], wall_s=round(time.perf_counter() - t0, 3))
    for p in fed.pools:
        p.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    for w in which:
        {"c2": c2, "c3": c3, "c4": c4, "c5": c5}[w]()
