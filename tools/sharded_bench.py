"""us/tick of a sharded pool in LAN steady state for several shard sizes (dev tool, torchrun)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consul_b200.pool import lan_config
from consul_b200.sharded import ShardedPool
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for per_mi in [int(x) for x in os.environ.get("GSIM_PER_MI", "2,8").split(",")]:
    n = per_mi * 1024 * 1024 * world
    for flags in [int(x) for x in os.environ.get("GSIM_FLAGS_LIST", "0,2").split(",")]:
        p = ShardedPool(lan_config(capacity=n, n_initial=n, seed=0x5EED0001, device=local, flags=flags))
        p.step(64)
        best = 1e9
        for _ in range(3):
            p.step(256)
            ms, nl = p.last_step_timing()
            best = min(best, ms * 1e3 / nl)
        t = torch.tensor([best], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world={world} per_gpu={per_mi}Mi flags={flags}: {t.item():.2f} us/tick -> {n / t.item() / 1e3:.1f} G node-ticks/s", flush=True)
        p.close()
dist.destroy_process_group()
