"""gloo worker for tests/test_sharded_cpu.py: the sharded-pool host logic (descriptor exchange,
controller protocol, per-tick barrier, member ranges) on the host emulation, world_size ranks,
against the same pool unsharded and against the oracle."""
import json
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from consul_b200 import _lib  # noqa: E402
from consul_b200.pool import (PRED_ALL_RUMORS_CONVERGED, PRED_CRASHED_ALL_DEAD, Pool,  # noqa: E402
                              consul_test_config, lan_config)
from consul_b200.sharded import ShardedPool  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
L = _lib.load(os.path.join(ROOT, "tests", "hostemu", "libgsim_hostemu.so"))


def script(p):
    out = {}
    x = p.member_add()
    out["joined"] = p.join(x, [5])
    slot = p.user_event(3, b"deploy", b"x" * 32, False)
    out["crashed"] = p.crash_fraction(50000, 1)
    p.step(37)
    y = p.member_add()
    p.join(y, [x, 7])
    p.leave(11)
    out["conv"] = p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 200, 4)
    out["dead_tick"] = p.run_until(PRED_CRASHED_ALL_DEAD, 0, 3000, 50)
    out["event"] = p.rumor_info(slot)
    st = p.stats()
    st.pop("active_rows")
    out["stats"] = st
    out["hash"] = ["%016x" % h for h in p.state_hash()]
    out["now"] = p.now
    out["nodes_seen_by_y"] = p.num_nodes(y)
    return out


N = 3 * 4096 * world - 100          # not a multiple of the shard size: the last rank is short
mk = lambda: lan_config(L, capacity=N + 8, n_initial=N, seed=0x5EED0009, packet_loss_ppm=30000)  # noqa: E731
sp = ShardedPool(mk(), L)
got = script(sp)
if rank == 0:
    keys = sp.column("key")          # bulk observation is served by rank 0
    assert int((keys & 3 == 2).sum()) == got["crashed"]
sp.close()
ok = True
if rank == 0:
    from oracle_binding import OraclePool
    for name, ref in (("unsharded", Pool(mk(), L)), ("oracle", OraclePool(mk()))):
        want = script(ref)
        for k in got:
            if got[k] != want[k]:
                ok = False
                print("MISMATCH vs", name, k, got[k], want[k], flush=True)
    print(json.dumps({"ok": ok, "world": world, "members": N, "hash": got["hash"][0]}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
