"""Steady-state window timing of one or more libgsim builds (dev tool): us per window launch and per tick."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consul_b200 import _lib
from consul_b200.pool import Pool, lan_config
for path in sys.argv[1:]:
    lib = _lib.load(path)
    for n, ticks in ((1_000_000, 2000), (16_777_216, 400), (67_108_864, 200)):
        p = Pool(lan_config(lib, capacity=n, n_initial=n, seed=0x5EED0001), lib)
        p.step(40)
        c0 = p.sched_counts()
        p.step(ticks)
        c1 = p.sched_counts()
        wl, wt, wms = (c1[k] - c0[k] for k in ("window_launches", "window_ticks", "window_ms"))
        print(f"{os.path.basename(path)} n={n}: {wl} window launches, {wt} ticks, {wms * 1e3 / max(wl, 1):.2f} us/launch, "
              f"{wms * 1e3 / max(wt, 1):.3f} us/tick, {n * wt / max(wms, 1e-9) / 1e6:.0f} G node-ticks/s, hash {p.state_hash()[0]:016x}", flush=True)
        p.close()
