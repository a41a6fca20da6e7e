"""BASELINE config 4 driver: 16 777 216 members, one UserEvent("deploy", 32-byte payload) fired at member 0,
run until every member has delivered it, then drain the retransmissions.  Prints one JSON line.

  1 GPU :  python tools/c4_event.py
  N GPUs:  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c4_event.py [--check]

With --check rank 0 replays the same script on ONE GPU afterwards and the 256-bit state digest, the
convergence tick and the counters must be identical ("identical state hash for every G")."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def script(p, drain):
    t0 = time.time()
    slot = p.user_event(0, b"deploy", bytes(32), False)
    from consul_b200.pool import PRED_RUMOR_CONVERGED
    t = p.run_until(PRED_RUMOR_CONVERGED, slot, 400, 4)
    ms_conv = p.last_step_timing()[0]
    p.step(drain)
    ms_drain = p.last_step_timing()[0]
    st = p.stats()
    info = p.rumor_info(slot)
    return {"ticks_to_convergence": int(t) + 1 if t != 0xFFFFFFFF else None, "now": p.now,
            "digest": ["%016x" % h for h in p.state_hash()], "heard": info["heard_count"],
            "rumors_accepted": st["rumors_accepted"], "rumors_sent": st["rumors_sent"], "gossip_packets": st["gossip_packets"],
            "kernel_ms": ms_conv + ms_drain, "wall_s": time.time() - t0, "sched": p.sched_counts()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=16 * 1024 * 1024)
    ap.add_argument("--drain", type=int, default=96)
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0x5EED0004)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import torch
    from consul_b200.pool import Pool, lan_config
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    cfg = lambda: lan_config(capacity=a.members, n_initial=a.members, seed=a.seed, device=local)
    if world > 1:
        import torch.distributed as dist
        from consul_b200.sharded import ShardedPool
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        p = ShardedPool(cfg())
    else:
        p = Pool(cfg())
    got = script(p, a.drain)
    p.close()
    if rank == 0:
        out = {"config": "C4: %d members, one UserEvent, LAN defaults" % a.members, "n_gpus": world, **got,
               "node_ticks_per_s_kernel": a.members * got["now"] / (got["kernel_ms"] / 1e3)}
        if a.check and world > 1:
            ref = Pool(cfg())
            want = script(ref, a.drain)
            ref.close()
            keys = ("ticks_to_convergence", "now", "digest", "heard", "rumors_accepted", "rumors_sent", "gossip_packets")
            out["digest_single_gpu"] = want["digest"]
            out["single_gpu_kernel_ms"] = want["kernel_ms"]
            out["parity_ok"] = all(got[k] == want[k] for k in keys)
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
