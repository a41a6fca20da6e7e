"""Join-cascade timing of one or more libgsim builds (dev tool): single-tick launches while the
joiner's alive broadcast and join intent run through a 1 M-member pool, then quiet windows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consul_b200 import _lib
from consul_b200.pool import Pool, lan_config
for path in sys.argv[1:]:
    lib = _lib.load(path)
    n = 1_000_000
    p = Pool(lan_config(lib, capacity=n + 16, n_initial=n, seed=0x5EED0001), lib)
    p.step(64)
    best = None
    for rep in range(4):
        c0 = p.sched_counts()
        x = p.member_add(); p.join(x, [0])
        p.step(2048)
        c1 = p.sched_counts()
        d = {k: c1[k] - c0[k] for k in c1}
        if best is None or d["tick_ms"] + d["window_ms"] < best["tick_ms"] + best["window_ms"]:
            best = d
    print(f"{os.path.basename(path)}: per 2048-tick step: {best['tick_launches']} single ticks {best['tick_ms']:.3f} ms "
          f"({best['tick_ms'] * 1e3 / max(best['tick_launches'], 1):.1f} us each), {best['window_launches']} windows "
          f"{best['window_ms']:.3f} ms ({best['window_ms'] * 1e3 / max(best['window_launches'], 1):.1f} us each), "
          f"hash {p.state_hash()[0]:016x}", flush=True)
    p.close()
