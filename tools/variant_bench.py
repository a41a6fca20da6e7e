"""Time several builds of libgsim (kernel variants) on the same workloads.  Dev tool."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consul_b200 import _lib
from consul_b200.pool import Pool, lan_config

def run(lib, n, ticks, join=False, crash=0, event=False):
    p = Pool(lan_config(lib, capacity=n + 1, n_initial=n, seed=0x5EED0001), lib)
    if join:
        x = p.member_add(); p.join(x, [0])
    if event:
        p.user_event(0, b"deploy", bytes(32), False)
    if crash:
        p.crash_fraction(crash, 0)
    p.step(64)
    best = 1e9
    for _ in range(3):
        p.step(ticks)
        ms, nl = p.last_step_timing()
        best = min(best, ms * 1e3 / nl)
    h = p.state_hash()[0]
    p.close()
    return best, h

for path in sys.argv[1:]:
    lib = _lib.load(path)
    out = []
    for (n, ticks, kw) in [(1_000_000, 1024, {}), (16_777_216, 256, {}), (67_108_864, 128, {}),
                           (1_000_000, 64, dict(event=True)), (4_000_000, 256, dict(crash=100000))]:
        us, h = run(lib, n, ticks, **kw)
        out.append(f"n={n} {kw or 'steady'}: {us:.2f} us/tick ({n / us / 1e3:.1f} G nt/s) hash {h:016x}")
    print(os.path.basename(path)); print("  " + "\n  ".join(out), flush=True)
