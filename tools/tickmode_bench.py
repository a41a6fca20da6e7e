"""Single-tick-mode timings (dev tool): LAN steady state with windows off at 1 M / 16 Mi members, and the
BASELINE config 3 crash wave at 4 M members (every tick is a single tick while suspicions run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consul_b200.pool import Pool, lan_config, FLAG_NO_WINDOWS
for n, ticks in ((1_000_000, 1024), (16_777_216, 256)):
    p = Pool(lan_config(capacity=n, n_initial=n, seed=0x5EED0001, flags=FLAG_NO_WINDOWS))
    p.step(64)
    c0 = p.sched_counts(); p.step(ticks); c1 = p.sched_counts()
    print(f"steady single ticks n={n}: {(c1['tick_ms'] - c0['tick_ms']) * 1e3 / ticks:.2f} us/tick, hash {p.state_hash()[0]:016x}", flush=True)
    p.close()
n = 4_000_000
p = Pool(lan_config(capacity=n, n_initial=n, seed=0x5EED0003))
p.crash_fraction(100000, 0)
p.step(64)
c0 = p.sched_counts(); p.step(512); c1 = p.sched_counts()
print(f"C3 crash wave n={n}: {(c1['tick_ms'] - c0['tick_ms']) * 1e3 / max(1, c1['tick_launches'] - c0['tick_launches']):.2f} us/tick over "
      f"{c1['tick_launches'] - c0['tick_launches']} single ticks, hash {p.state_hash()[0]:016x}", flush=True)
