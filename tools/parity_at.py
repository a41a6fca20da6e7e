"""bench.py's parity replay at an arbitrary size on one GPU (dev tool): python tools/parity_at.py 8388352 8388608"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from consul_b200.pool import Pool, lan_config
from oracle_binding import OraclePool, oracle_config
n, cap = int(sys.argv[1]), int(sys.argv[2])
got = bench.parity_script(Pool(lan_config(capacity=cap, n_initial=n, seed=bench.SEED)))
want = bench.parity_script(OraclePool(oracle_config("lan", capacity=cap, n_initial=n, seed=bench.SEED),
                                      threads=max(1, bench.host_cores()["physical"] // 2)))
print(json.dumps({"members": n, "gpu": got, "oracle": want, "parity_ok": got == want}), flush=True)
