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
                              PRED_RUMOR_CONVERGED, consul_test_config, lan_config, wan_config)
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
    out["update_slot"] = p.member_update(13, 90)              # SetTags: every key replica gets the new incarnation
    out["conv"] = p.run_until(PRED_ALL_RUMORS_CONVERGED, 0, 200, 4)
    out["dead_tick"] = p.run_until(PRED_CRASHED_ALL_DEAD, 0, 3000, 50)
    out["event"] = p.rumor_info(slot)
    p.step(60)                                                # the reaper (2 s / every 1 s) erases the dead
    st = p.stats()
    st.pop("active_rows")
    out["stats"] = st
    out["hash"] = ["%016x" % h for h in p.state_hash()]
    out["now"] = p.now
    out["nodes_seen_by_y"] = p.num_nodes(y)
    return out


def wan_script(p):
    """WAN latency pool (BASELINE config 5): deeper mailbox ring, deliveries that cross shards
    several ticks after they were sent."""
    from consul_b200.wan import c5_latency_matrix
    out = {}
    p.latency_set(c5_latency_matrix(64))
    p.member_watch(130, True)
    slot = p.user_event(1, b"deploy", b"v2", False)
    p.crash_many([200, 9000])
    p.step(5)
    out["hash5"] = ["%016x" % h for h in p.state_hash()]       # packets in flight in the ring
    out["conv"] = p.run_until(PRED_RUMOR_CONVERGED, slot, 400, 3)
    out["inject"] = p.rumor_inject(slot, 7)                     # already heard: not accepted
    p.step(60)
    out["event"] = p.rumor_info(slot)
    st = p.stats()
    st.pop("active_rows")
    out["stats"] = st
    out["hash"] = ["%016x" % h for h in p.state_hash()]
    out["events"] = [(e.tick, e.type, e.subject, e.observer) for e in p.poll_events()] if getattr(p, "rank", 0) == 0 else None
    return out


N = 3 * 4096 * world - 100          # not a multiple of the shard size: the last rank is short
mk = lambda: lan_config(L, capacity=N + 8, n_initial=N, seed=0x5EED0009, packet_loss_ppm=30000,  # noqa: E731
                        flags=32, push_pull_interval_ns=10**9,   # periodic push-pull on (every ~100 ticks)
                        reconnect_timeout_ns=2 * 10**9, tombstone_timeout_ns=2 * 10**9, reap_interval_ns=10**9)
sp = ShardedPool(mk(), L)
got = script(sp)
if rank == 0:
    keys = sp.column("key")          # bulk observation is served by rank 0
    assert int((keys & 3 == 2).sum()) == got["stats"]["n_crashed"] < got["crashed"]   # the reaper erased some
sp.close()
mkw = lambda: wan_config(L, capacity=N, n_initial=N, seed=0x5EED000A, mailbox_depth=8)  # noqa: E731
spw = ShardedPool(mkw(), L)
gotw = wan_script(spw)
spw.close()


def federation_script(make_pool):
    """Two co-sharded WAN pools with bridge members (BASELINE config 5 in small)."""
    from consul_b200.wan import WanFederation
    pa, pb = make_pool(0x5EED000B), make_pool(0x5EED000C)
    fed = WanFederation(pa, pb, n_dcs=16, bridges_per_dc=2, n_members=N)
    fed.fire(0, 1, b"deploy", b"v3")
    t = fed.run_until_converged(b"deploy", b"v3", 400)
    out = {"t": t, "forwarded": fed.forwarded, "slots": fed.slots[(b"deploy", b"v3")],
           "hash": [["%016x" % h for h in p.state_hash()] for p in fed.pools],
           "info": [p.rumor_info(s) for p, s in zip(fed.pools, fed.slots[(b"deploy", b"v3")])]}
    for p in fed.pools:
        getattr(p, "close", lambda: None)()
    return out


# randomised operation sequences through the controller protocol (every rank issues every call and
# keeps its own oracle; digest and counters are compared after every operation)
import fuzz_ops  # noqa: E402
from oracle_binding import OraclePool as _Oracle  # noqa: E402
fuzz_ok = True
for fseed in (3, 17, 251, 404):
    try:
        fuzz_ops.run_sequence(lambda cfg: [ShardedPool(cfg, L), _Oracle(cfg)], L, fseed, n_ops=40, columns=False,
                              single_gpu_features=False)
    except AssertionError as e:
        fuzz_ok = False
        print("FUZZ MISMATCH rank", rank, "seed", fseed, str(e)[:800], flush=True)
for fseed in (3001, 3004, 3013):   # calm pools: quiet windows (closed form included) with one barrier per launch
    try:
        fuzz_ops.run_sequence(lambda cfg: [ShardedPool(cfg, L), _Oracle(cfg)], L, fseed, n_ops=30, columns=False,
                              single_gpu_features=False, calm=True)
    except AssertionError as e:
        fuzz_ok = False
        print("CALM FUZZ MISMATCH rank", rank, "seed", fseed, str(e)[:800], flush=True)

mkf = lambda seed: wan_config(L, capacity=N, n_initial=N, seed=seed, mailbox_depth=8)  # noqa: E731
gotf = federation_script(lambda seed: ShardedPool(mkf(seed), L))
ok = fuzz_ok
if rank == 0:
    from oracle_binding import OraclePool
    for name, ref in (("unsharded", Pool(mk(), L)), ("oracle", OraclePool(mk()))):
        want = script(ref)
        for k in got:
            if got[k] != want[k]:
                ok = False
                print("MISMATCH vs", name, k, got[k], want[k], flush=True)
    for name, ref in (("unsharded", Pool(mkw(), L)), ("oracle", OraclePool(mkw()))):
        want = wan_script(ref)
        for k in gotw:
            if gotw[k] != want[k]:
                ok = False
                print("WAN MISMATCH vs", name, k, gotw[k], want[k], flush=True)
    wantf = federation_script(lambda seed: OraclePool(mkf(seed)))
    for k in gotf:
        if gotf[k] != wantf[k]:
            ok = False
            print("FEDERATION MISMATCH vs oracle", k, gotf[k], wantf[k], flush=True)
    print(json.dumps({"ok": ok, "world": world, "members": N, "hash": got["hash"][0]}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
