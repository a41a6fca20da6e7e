"""torchrun worker for tests/test_gpu_sharded.py: the same script on a pool sharded over all
ranks and on a single-GPU pool (rank 0); SURVEY §8e invariant: identical state for every G."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from consul_b200.pool import PRED_ALL_RUMORS_CONVERGED, PRED_CRASHED_ALL_DEAD, Pool, lan_config  # noqa: E402
from consul_b200.sharded import ShardedPool  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
per = int(os.environ.get("GSIM_ROWS_PER_RANK", 2 * 1024 * 1024))
N = per * world
SEED = 0x5EED0004


def script(p):
    out = {}
    x = p.member_add()
    out["joined"] = p.join(x, [5])
    slot = p.user_event(3, b"deploy", b"x" * 32, False)
    out["crashed"] = p.crash_fraction(20000, 1)
    t0 = time.perf_counter()
    if os.environ.get("GSIM_SHARD_STEP1"):      # debug: a host-level barrier around every tick
        for _ in range(192):
            p.step(1)
    else:
        p.step(192)
    out["wall_s_192"] = time.perf_counter() - t0
    out["us_per_tick"] = p.last_step_timing()[0] * 1e3 / 192
    out["conv"] = p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 64, 8)
    out["dead_tick"] = p.run_until(PRED_CRASHED_ALL_DEAD, 0, 640, 64)
    out["event"] = p.rumor_info(slot) if (slot is not None) else None
    st = p.stats()
    st.pop("active_rows")
    out["stats"] = st
    out["hash"] = ["%016x" % h for h in p.state_hash()]
    out["now"] = p.now
    return out


cfg = lan_config(capacity=N, n_initial=N - 8, seed=SEED, device=local, flags=int(os.environ.get("GSIM_FLAGS", "0")))
sp = ShardedPool(cfg)
got = script(sp)
sp.close()
ok = True
if rank == 0:
    ref = Pool(lan_config(capacity=N, n_initial=N - 8, seed=SEED, device=local))
    want = script(ref)
    ref.close()
    for k in ("joined", "crashed", "conv", "dead_tick", "event", "stats", "hash", "now"):
        if got[k] != want[k]:
            ok = False
            if isinstance(got[k], dict):
                print("MISMATCH", k, {f: (got[k][f], want[k][f]) for f in got[k] if got[k][f] != want[k][f]}, flush=True)
            else:
                print("MISMATCH", k, got[k], want[k], flush=True)
    print(json.dumps({"ok": ok, "world": world, "members": N, "sharded_us_per_tick": got["us_per_tick"],
                      "single_us_per_tick": want["us_per_tick"], "dead_tick": got["dead_tick"],
                      "hash": got["hash"][0]}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
