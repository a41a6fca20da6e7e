"""Members() of a 1 M-member pool on one GPU: wall time per call for 1, 2, 4, 8 host threads (dev tool)."""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    from consul_b200._lib import GsimMember
    from consul_b200.pool import Pool, lan_config
    n = 1_000_000
    p = Pool(lan_config(capacity=n + 64, n_initial=n, seed=1))
    buf = (GsimMember * (n + 64))()
    k = C.c_size_t()
    ts = []
    for _ in range(12):
        t = time.perf_counter()
        rc = p.lib.gsim_members(p.h, 0, buf, n + 64, C.byref(k))
        ts.append((time.perf_counter() - t) * 1e3)
    print(sys.argv[1], "threads:", rc, k.value, "ms best3", [round(x, 3) for x in sorted(ts)[:3]], flush=True)
else:
    for t in ("1", "2", "4", "8", "16"):
        subprocess.run([sys.executable, __file__, t], env=dict(os.environ, GSIM_MEMBERS_THREADS=t))
