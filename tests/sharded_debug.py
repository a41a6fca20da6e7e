"""Debug aid (2 GPUs): find the first tick at which the sharded pool diverges from the single-GPU
pool and print which columns/rows differ."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from consul_b200.pool import Pool, lan_config  # noqa: E402
from consul_b200.sharded import ShardedPool  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
per = 2 * 1024 * 1024
N = per * world
T = int(os.environ.get("GSIM_DEBUG_TICKS", 80))
CHUNK = int(os.environ.get("GSIM_DEBUG_CHUNK", 1))
mk = lambda: lan_config(capacity=N, n_initial=N - 8, seed=0x5EED0004, device=local)  # noqa: E731
COLS = ["key", "meta", "due", "cursor", "pass", "probe_tgt", "probe_inc", "sus_start", "heard", "queued", "inbox",
        "change_tick", "ltime_member", "ltime_event"]


def prep(p):
    x = p.member_add()
    p.join(x, [5])
    p.user_event(3, b"deploy", b"x" * 32, False)
    p.crash_fraction(20000, 1)


sp = ShardedPool(mk())
prep(sp)
ref = None
if rank == 0:
    ref = Pool(mk())
    prep(ref)
first = None
for t in range(0, T, CHUNK):
    sp.step(CHUNK)
    hs = sp.state_hash()
    ss = sp.stats()
    if rank == 0:
        ref.step(CHUNK)
        hr = ref.state_hash()
        sr = ref.stats()
        bad = hs != hr
        flag = torch.tensor([1 if bad else 0], device="cuda")
    else:
        flag = torch.tensor([0], device="cuda")
    dist.broadcast(flag, 0)
    if flag.item():
        first = t + 1
        if rank == 0:
            print("FIRST DIVERGENCE after tick", t, "now", sp.now, flush=True)
            print("stats diff", {k: (ss[k], sr[k]) for k in ss if ss[k] != sr[k]}, flush=True)
            for c in COLS:
                a, b = sp.column(c)[:N], ref.column(c)[:N]
                idx = np.nonzero(a != b)[0]
                if len(idx):
                    print(f"column {c}: {len(idx)} rows differ, first {idx[:8].tolist()} "
                          f"sharded {a[idx[:8]].tolist()} single {b[idx[:8]].tolist()}", flush=True)
        break
if first is None and rank == 0:
    print("NO DIVERGENCE in", T, "ticks", flush=True)
sp.close()
dist.barrier()
dist.destroy_process_group()
