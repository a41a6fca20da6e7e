"""Per-rank timings of a sharded pool (dev tool, torchrun): quiet windows, single ticks in steady state,
and a join cascade, 1 Mi members per GPU.  GSIM_FLAGS adds pool flags (8 = lean fence)."""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consul_b200.pool import lan_config, Pool, FLAG_NO_WINDOWS
from consul_b200.sharded import ShardedPool
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
per = int(os.environ.get("GSIM_PER", 1024 * 1024))
extra = int(os.environ.get("GSIM_FLAGS", "0"))
n = per * world - 64


def delta(p, fn):
    c0 = p.sched_counts(); fn(); c1 = p.sched_counts()
    return {k: c1[k] - c0[k] for k in c1}


out = {}
for name, flags in (("windows", extra), ("single_ticks", extra | FLAG_NO_WINDOWS)):
    p = ShardedPool(lan_config(capacity=per * world, n_initial=n, seed=0x5EED0001, device=local, flags=flags))
    p.step(64)
    d = delta(p, lambda: p.step(1280))
    out[name] = {"window_us_per_launch": d["window_ms"] * 1e3 / max(1, d["window_launches"]), "window_launches": d["window_launches"],
                 "tick_us": d["tick_ms"] * 1e3 / max(1, d["tick_launches"]), "ticks": d["tick_launches"]}
    if name == "windows":
        def cascade():
            x = p.member_add(); p.join(x, [0]); p.step(2048)
        cascade()
        d = delta(p, cascade)
        # where a bench step's wall time goes: each API call timed on the host (every rank issues every call)
        import time
        wall = {"member_add": 0.0, "join": 0.0, "step": 0.0, "stats": 0.0}
        REP = 6
        for _ in range(REP):
            torch.cuda.synchronize(); t0 = time.perf_counter(); x = p.member_add()
            t1 = time.perf_counter(); p.join(x, [0])
            t2 = time.perf_counter(); p.step(2048)
            t3 = time.perf_counter(); p.stats()
            t4 = time.perf_counter()
            for k, v in zip(wall, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                wall[k] += v * 1e3 / REP
        out["wall_ms_per_call"] = {k: round(v, 3) for k, v in wall.items()}
        out["cascade"] = {"tick_us": d["tick_ms"] * 1e3 / max(1, d["tick_launches"]), "ticks": d["tick_launches"],
                          "window_us_per_launch": d["window_ms"] * 1e3 / max(1, d["window_launches"]), "window_launches": d["window_launches"]}
    p.close()
if rank == 0 and os.environ.get("GSIM_SINGLE", "1") == "1":      # the same shard size on one GPU, unsharded
    p = Pool(lan_config(capacity=per, n_initial=per - 64, seed=0x5EED0001, device=local))
    p.step(64)
    d = delta(p, lambda: p.step(1280))
    out["one_gpu_unsharded"] = {"window_us_per_launch": d["window_ms"] * 1e3 / max(1, d["window_launches"]), "window_launches": d["window_launches"]}
    p.close()
print(json.dumps({"rank": rank, "world": world, "per_gpu": per, **out}), flush=True)
dist.barrier()
dist.destroy_process_group()
