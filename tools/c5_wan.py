"""BASELINE config 5 driver: two WAN pools with the asymmetric 64-datacenter latency matrix and 5
bridge members per datacenter, one event fired at A/DC0, run until every member of both pools has
delivered it.  Prints one JSON line (ticks to convergence, bridge re-fires, kernel time).

  1 GPU :  python tools/c5_wan.py --members 8388608
  N GPUs:  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c5_wan.py --members M

Both pools are sharded over all N ranks (co-sharded).  A sharded pool holds a multiple of 1 Mi
members per rank (2 MB mapping granularity over the narrowest, 2-byte column), so the 8 Mi pools
of config 5 shard evenly over 1, 2, 4 or 8 GPUs.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=8 * 1024 * 1024, help="members per pool")
    ap.add_argument("--dcs", type=int, default=64)
    ap.add_argument("--bridges", type=int, default=5)
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0x5EED0051)
    ap.add_argument("--max-ticks", type=int, default=600)
    ap.add_argument("--check", action="store_true", help="rank 0 replays the run on ONE GPU and compares digests")
    a = ap.parse_args()

    import torch
    from consul_b200.pool import Pool, wan_config
    from consul_b200.wan import WanFederation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        from consul_b200.sharded import ShardedPool
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        mk = lambda seed: ShardedPool(wan_config(capacity=a.members, n_initial=a.members, seed=seed,
                                                 mailbox_depth=8, device=local))
    else:
        mk = lambda seed: Pool(wan_config(capacity=a.members, n_initial=a.members, seed=seed, mailbox_depth=8))
    t0 = time.time()
    fed = WanFederation(mk(a.seed), mk(a.seed + 1), n_dcs=a.dcs, bridges_per_dc=a.bridges, n_members=a.members)
    setup_s = time.time() - t0
    name, payload = b"deploy", b"x" * 32
    fed.fire(0, a.bridges + 2, name, payload)            # a non-bridge member of A/DC0
    kernel_ms = 0.0
    t0 = time.time()
    ticks = None
    for _ in range(a.max_ticks):
        if fed.converged((name, payload)):
            ticks = max(p.rumor_info(s)["converged_tick"] for p, s in zip(fed.pools, fed.slots[(name, payload)]))
            break
        fed.step(1)
        kernel_ms += sum(p.last_step_timing()[0] for p in fed.pools)
    wall_s = time.time() - t0
    stats = [p.stats() for p in fed.pools]               # collective on sharded pools: every rank calls
    digests = [["%016x" % h for h in p.state_hash()] for p in fed.pools]
    now = fed.pools[0].now
    check = {}
    if a.check and world > 1:
        for p in fed.pools:
            p.close()
        if rank == 0:                                    # the same federation on one GPU
            one = lambda seed: Pool(wan_config(capacity=a.members, n_initial=a.members, seed=seed, mailbox_depth=8, device=local))
            ref = WanFederation(one(a.seed), one(a.seed + 1), n_dcs=a.dcs, bridges_per_dc=a.bridges, n_members=a.members)
            ref.fire(0, a.bridges + 2, name, payload)
            t_ref = ref.run_until_converged(name, payload, a.max_ticks)
            d_ref = [["%016x" % h for h in p.state_hash()] for p in ref.pools]
            check = {"digest_single_gpu": d_ref, "ticks_single_gpu": t_ref,
                     "parity_ok": d_ref == digests and t_ref == ticks and ref.forwarded == fed.forwarded}
            for p in ref.pools:
                p.close()
    if rank == 0:
        print(json.dumps({**check, "digest": digests,
            "config": "C5: two WAN pools, %d members each, %d DCs, %d bridges/DC, L[a][b]=1+((7a+13b) mod 5)"
                      % (a.members, a.dcs, a.bridges),
            "n_gpus": world, "ticks_to_convergence": ticks, "ticks_run": now,
            "bridge_refires": fed.forwarded, "refires_into": fed.forwarded_into,
            "suspects": [s["suspects"] for s in stats], "probes": [s["probes"] for s in stats],
            "kernel_ms": kernel_ms, "wall_s": wall_s, "setup_s": setup_s,
            "node_ticks_per_s_kernel": 2 * a.members * now / (kernel_ms / 1e3) if kernel_ms else None,
        }), flush=True)
    if not (a.check and world > 1):
        for p in fed.pools:
            p.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
