"""Random operation sequences replayed on two pools (kernel body / CUDA vs oracle) and compared after
every step: the scenario tests cover the paths somebody thought of, this covers the others."""
from __future__ import annotations

import random

import numpy as np

from consul_b200.pool import FLAG_COORDINATES, FLAG_PUSH_PULL, consul_test_config, lan_config, wan_config
from parity import check_invariants, compare_pools

MS = 1_000_000


def random_config(rng: random.Random, lib, calm: bool = False):
    """calm = a pool that can become quiet and pristine (no loss, no periodic tickers besides the probe):
    the sequences then spend most of their ticks in quiet windows, closed form included."""
    preset = rng.choice(["lan", "lan", "wan", "test"])
    n0 = rng.choice([0, 1, 2, 3, 5, 40, 130, 300, 700])
    kw = dict(capacity=n0 + 24, n_initial=n0, seed=rng.getrandbits(48), flags=1)
    if calm:
        n0 = rng.choice([3, 17, 40, 130, 300, 700])
        kw.update(capacity=n0 + 24, n_initial=n0)
        if rng.random() < 0.3:
            kw["retransmit_mult"] = rng.choice([1, 2])
        if rng.random() < 0.3:
            kw["gossip_nodes"] = 1
        if rng.random() < 0.3:
            kw.update(reap_interval_ns=rng.choice([100, 500]) * MS, reconnect_timeout_ns=rng.choice([200, 2000]) * MS,
                      tombstone_timeout_ns=rng.choice([200, 2000]) * MS)
        if n0 < 128 or rng.random() < 0.3:
            kw["phase_group"] = 1
        fn = {"lan": lan_config, "wan": wan_config, "test": consul_test_config}[preset]
        return fn(lib, **kw), None
    if rng.random() < 0.5:
        kw["packet_loss_ppm"] = rng.choice([20000, 150000, 450000])
    if rng.random() < 0.25:
        kw["disable_tcp_pings"] = 1
    if rng.random() < 0.3:
        kw["udp_buffer_size"] = rng.choice([90, 140, 400])
    if rng.random() < 0.3:
        kw["retransmit_mult"] = rng.choice([1, 2])
    if rng.random() < 0.3:
        kw["event_buffer"] = rng.choice([2, 8])
    if rng.random() < 0.3:
        kw["flags"] |= FLAG_PUSH_PULL
        kw["push_pull_interval_ns"] = rng.choice([300, 1000, 4000]) * MS
    if rng.random() < 0.2:
        kw["flags"] |= FLAG_COORDINATES                      # the digest folds every coordinate bit for bit
    if rng.random() < 0.3:
        kw.update(reap_interval_ns=rng.choice([100, 500]) * MS, reconnect_timeout_ns=rng.choice([200, 2000]) * MS,
                  tombstone_timeout_ns=rng.choice([200, 2000]) * MS)
    if n0 < 128 or rng.random() < 0.3:
        kw["phase_group"] = 1
    latency = None
    if rng.random() < 0.35:
        kw["mailbox_depth"] = 8
        n_dcs = rng.choice([2, 5, 16])
        worst = rng.choice([2, 5, 7])
        a = np.arange(n_dcs)[:, None]
        b = np.arange(n_dcs)[None, :]
        latency = (1 + (3 * a + 5 * b) % worst).astype(np.uint8)
        latency[np.arange(n_dcs), np.arange(n_dcs)] = 1
    fn = {"lan": lan_config, "wan": wan_config, "test": consul_test_config}[preset]
    return fn(lib, **kw), latency


def random_graph(rng: random.Random, n: int):
    """A sparse static topology: every member knows itself, a window of neighbours and a few
    random others (directed; some members may end up knowing nobody else)."""
    rows = []
    for i in range(n):
        row = {i} | {(i + d) % n for d in range(1, rng.choice([1, 2, 9]))} | {rng.randrange(n) for _ in range(rng.choice([0, 2, 6]))}
        rows.append(sorted(row))
    rp = np.zeros(n + 1, dtype=np.uint32)
    rp[1:] = np.cumsum([len(r) for r in rows])
    return rp, np.concatenate([np.array(r, dtype=np.uint32) for r in rows])


def run_sequence(make, lib, seed: int, n_ops: int = 60, columns: bool = True, single_gpu_features: bool = True,
                 calm: bool = False):
    rng = random.Random(seed)
    cfg, latency = random_config(rng, lib, calm)
    if not single_gpu_features:                                # sharded pools: no network coordinates yet
        cfg.flags &= ~FLAG_COORDINATES
    pools = make(cfg)
    if latency is not None:
        for p in pools:
            p.latency_set(latency)
    want_graph = cfg.n_initial >= 5 and rng.random() < 0.2 and not calm   # static CSR topology (member_add then fails alike)
    if want_graph and getattr(pools[0], "world", 1) == 1:      # (peer graphs are single-GPU for now)
        rp, ci = random_graph(rng, cfg.n_initial)
        for p in pools:
            p.graph_set(rp, ci)
    log = []
    inv = None

    def both(fn, what):
        out = []
        for p in pools:
            try:
                out.append(("ok", fn(p)))
            except Exception as e:                         # both must fail alike
                out.append(("err", getattr(e, "code", None)))
        log.append((what, out[0]))
        kinds = [o[0] for o in out]
        assert kinds[0] == kinds[1], (seed, log[-5:], out)
        if kinds[0] == "ok":
            assert out[0][1] == out[1][1], (seed, log[-5:], out)
        return out[0]

    for step in range(n_ops):
        n = pools[0].stats()["n_members"]
        r = rng.random()
        if calm and rng.random() < 0.6:
            r = 0.99                                           # mostly time passing
        elif calm and 0.30 <= r < 0.45:
            r = 0.05                                           # rather a joiner than a crash / leave: stays pristine
        pick = (lambda: rng.randrange(n)) if n else (lambda: 0)
        if r < 0.10:
            both(lambda p: p.member_add(watched=rng_bool(seed, step)), "add")
        elif r < 0.22 and n:
            a, b = pick(), pick()
            both(lambda p: p.join(a, [b, b], rng_bool(seed, step + 1)), f"join {a}->{b}")
        elif r < 0.30 and n:
            m = pick()
            both(lambda p: p.user_event(m, b"e%d" % step, b"x" * (step % 40), False), f"event {m}")
        elif r < 0.36 and n:
            m = pick()
            both(lambda p: p.crash(m), f"crash {m}")
        elif r < 0.41 and n:
            m = pick()
            both(lambda p: p.leave(m), f"leave {m}")
        elif r < 0.45 and n:
            a, b, pr = pick(), pick(), rng.random() < 0.5
            both(lambda p: p.force_leave(a, b, pr), f"force_leave {b} prune={pr}")
        elif r < 0.50 and n:
            m = pick()
            both(lambda p: p.member_update(m, 30 + step), f"update {m}")
        elif r < 0.54 and n:
            m, slot = pick(), rng.randrange(30)
            both(lambda p: p.rumor_inject(slot, m), f"inject {slot}->{m}")
        elif r < 0.57 and n:
            m = pick()
            both(lambda p: p.member_watch(m, True), f"watch {m}")
        elif r < 0.60 and n:
            m, t = pick(), rng.choice([0, 150, 900]) * MS
            both(lambda p: p.member_reconnect_timeout_set(m, t), f"rc_tm {m}")
        elif r < 0.63:
            slot = rng.randrange(30)
            both(lambda p: p.rumor_retire(slot), f"retire {slot}")
        elif r < 0.645 and hasattr(pools[0], "snapshot") and getattr(pools[0], "world", 1) == 1:
            ev = [sorted((e.tick, e.type, e.subject, e.observer, e.ltime) for e in p.poll_events()) for p in pools]
            assert ev[0] == ev[1], (seed, "event logs differ")     # (a restore starts with an empty event log)
            blob = pools[0].snapshot()                     # checkpoint / resume must be invisible
            pools[0].step(3)
            pools[0].restore(blob)
            log.append(("snapshot+3 ticks+restore", None))
        elif r < 0.66 and n > 10 and cfg.capacity > n:
            ppm = rng.choice([50000, 300000])
            both(lambda p: p.crash_fraction(ppm, step), f"crash_fraction {ppm}")
        else:
            k = rng.choice([1, 1, 2, 3, 7, 20, 64, 65, 130])
            if calm:                                       # long quiet stretches: ring passes end, own entries come up
                k = rng.choice([1, 3, 20, 130, 400, 1300, 2560, 2561, 6000])
            for p in pools:
                p.step(k)
            log.append((f"step {k}", None))
        try:
            compare_pools(pools[0], pools[1], f"seed {seed} after op {step}: {log[-1][0]}", columns=columns)
            if columns and step % 4 == 0:
                inv = check_invariants(pools[0], inv, f"seed {seed} after op {step}")
        except AssertionError as e:
            raise AssertionError(f"{e}\nlast ops: {log[-8:]}") from None
    if getattr(pools[0], "rank", 0) == 0:                      # a sharded pool's log is served by rank 0
        ev = [sorted((e.tick, e.type, e.subject, e.observer, e.ltime) for e in p.poll_events()) for p in pools]
        assert ev[0] == ev[1], (seed, "event logs differ")
    for p in pools:
        getattr(p, "close", lambda: None)()
    return len(log)


def rng_bool(seed, salt):
    return random.Random(seed * 1000003 + salt).random() < 0.5
