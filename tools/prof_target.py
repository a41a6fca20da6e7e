"""Small deterministic workload for ncu captures: N members, optional join cascade / crash."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consul_b200.pool import Pool, lan_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--members", type=int, default=1_000_000)
ap.add_argument("--ticks", type=int, default=320)
ap.add_argument("--join", action="store_true")
ap.add_argument("--event", action="store_true")
ap.add_argument("--crash-ppm", type=int, default=0)
ap.add_argument("--nograph", action="store_true")
a = ap.parse_args()
lib = None
if os.environ.get("GSIM_LIB"):                      # a kernel variant built by tools/build_variants.sh
    from consul_b200 import _lib
    lib = _lib.load(os.environ["GSIM_LIB"])
p = Pool(lan_config(lib, capacity=a.members + 1, n_initial=a.members, seed=0x5EED0001, flags=2 if a.nograph else 0), lib)
if a.join:
    x = p.member_add()
    p.join(x, [0])
if a.event:
    p.user_event(0, b"deploy", bytes(32), False)
if a.crash_ppm:
    p.crash_fraction(a.crash_ppm, 0)
p.step(a.ticks)
ms, n = p.last_step_timing()
print(f"members={a.members} ticks={a.ticks} kernel_ms={ms:.3f} launches={n} sched={p.sched_counts()}")
