"""Regenerates tests/golden/scenarios.json from the oracle (oracle/liboracle.so).

The Go modules that implement this path are not in /root/reference and there is no Go
toolchain, so there are no reference-produced vectors to commit.  These
fixtures pin the oracle's own behaviour: any change of semantics shows up as a diff here, and
the GPU tests compare the CUDA path with the same numbers.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

CASES = [
    # name, preset, n_initial, seed, script
    ("c1_three_node_join_crash", "test", 0, 1, [("add", 3), ("join", 1, 0), ("join", 2, 0), ("until", 2, 400),
                                                 ("step", 40), ("crash", 2), ("until", 3, 2000), ("step", 10)]),
    ("c2_join_4096", "lan", 4096, 0x5EED0001, [("add", 1), ("join", 4096, 0), ("until", 2, 600), ("step", 64)]),
    ("c2_join_100k", "lan", 100_000, 0x5EED0002, [("add", 1), ("join", 100_000, 0), ("until", 2, 600), ("step", 64)]),
    ("c3_crash_20k", "lan", 20_000, 0x5EED0003, [("crashf", 100000, 0), ("until", 3, 6000), ("step", 16)]),
    ("c4_event_50k", "lan", 50_000, 0x5EED0001, [("event", 0, "deploy", 32), ("until", 2, 600), ("step", 100)]),
    ("wan_crash_5k", "wan", 5_000, 7, [("crashf", 50000, 1), ("until", 3, 8000), ("step", 8)]),
    ("lossy_2k", "lan", 2_000, 11, [("loss", 150000), ("add", 1), ("join", 2000, 3), ("event", 5, "e1", 7),
                                   ("crash", 10), ("crash", 11), ("step", 600)]),
    ("leave_300", "lan", 300, 5, [("leave", 7), ("step", 150)]),
]

# Features added after round 1's GPU verification (WAN latency ring, push-pull, reaper, SetTags):
# their fixtures live in scenarios_ext.json so that scenarios.json stays byte-identical.
MS, SEC = 10**6, 10**9
CASES_EXT = [
    ("wan_c5_latency_event_16k", "wan", 16_384 + 77, 0x5EED0005,
     [("cfg", dict(mailbox_depth=8)), ("latency", "c5", 64), ("event", 0, "deploy", 32), ("until", 2, 800), ("step", 120)]),
    ("wan_slow_links_lossy_3k", "wan", 3_000, 21,
     [("cfg", dict(mailbox_depth=8, packet_loss_ppm=200000, disable_tcp_pings=1)), ("latency", "slow", 16, 7),
      ("add", 1), ("join", 3000, 3), ("event", 5, "e1", 7), ("crash", 10), ("crash", 300), ("crash", 1200), ("step", 700)]),
    ("pushpull_stranded_3k", "lan", 3_000, 0x5EED00AA,
     [("cfg", dict(flags=32, push_pull_interval_ns=1 * SEC, packet_loss_ppm=500000, retransmit_mult=1)),
      ("event", 17, "deploy", 2), ("add", 1), ("join", 3000, 3), ("crash", 100), ("crash", 200), ("step", 900)]),
    ("lan_reap_3", "test", 0, 1,
     [("cfg", dict(flags=1, reconnect_timeout_ns=250 * MS, tombstone_timeout_ns=250 * MS, reap_interval_ns=300 * MS)),
      ("add", 3), ("join", 1, 0), ("join", 2, 0), ("until", 2, 400), ("crash", 1), ("until", 3, 2000), ("step", 20),
      ("leave", 2), ("step", 120)]),
    ("set_tags_500", "lan", 500, 41,
     [("cfg", dict(packet_loss_ppm=200000)), ("step", 7), ("update", 5, 120), ("until", 2, 600), ("step", 50),
      ("update", 5, 0), ("step", 60)]),
]


def latency_matrix(kind, n_dcs, worst=5):
    import numpy as np
    a = np.arange(n_dcs)[:, None]
    b = np.arange(n_dcs)[None, :]
    m = 1 + ((7 * a + 13 * b) % 5 if kind == "c5" else (3 * a + 5 * b) % worst)
    m[np.arange(n_dcs), np.arange(n_dcs)] = 1
    return m.astype(np.uint8)


def run_case(make_pool, cfg_fns, case):
    name, preset, n0, seed, script = case
    extra = {}
    for op in script:
        if op[0] == "loss":
            extra["packet_loss_ppm"] = op[1]
        if op[0] == "cfg":
            extra.update(op[1])
    cap = n0 + 8
    kw = dict(capacity=cap, n_initial=n0, seed=seed, **extra)
    if preset == "test":
        kw["phase_group"] = 1
    pool = make_pool(cfg_fns[preset](**kw))
    untils = []
    for op in script:
        if op[0] == "add":
            for _ in range(op[1]):
                pool.member_add()
        elif op[0] == "join":
            pool.join(op[1], [op[2]])
        elif op[0] == "until":
            untils.append(pool.run_until(op[1], 0, op[2], 1))
        elif op[0] == "step":
            pool.step(op[1])
        elif op[0] == "crash":
            pool.crash(op[1])
        elif op[0] == "crashf":
            untils.append(pool.crash_fraction(op[1], op[2]))
        elif op[0] == "event":
            pool.user_event(op[1], op[2].encode(), b"x" * op[3], False)
        elif op[0] == "leave":
            pool.leave(op[1])
        elif op[0] == "latency":
            pool.latency_set(latency_matrix(op[1], op[2], op[3] if len(op) > 3 else 5))
        elif op[0] == "update":
            pool.member_update(op[1], op[2])
    st = pool.stats()
    st.pop("active_rows", None)
    return {"name": name, "tick": pool.now, "results": untils, "state_hash": [f"{h:016x}" for h in pool.state_hash()],
            "stats": {k: v for k, v in st.items() if k != "events_dropped"}}


def config_fns(lib=None):
    from consul_b200.pool import consul_test_config, lan_config, wan_config
    return {"lan": lambda **kw: lan_config(lib, **kw), "wan": lambda **kw: wan_config(lib, **kw),
            "test": lambda **kw: consul_test_config(lib, **kw)}


if __name__ == "__main__":
    from consul_b200 import _lib
    from oracle_binding import OraclePool
    fns = config_fns(_lib.lib())
    out = [run_case(lambda cfg: OraclePool(cfg, threads=1), fns, c) for c in CASES]
    with open(os.path.join(HERE, "scenarios.json"), "w") as f:
        json.dump(out, f, indent=1)
    for o in out:
        print(o["name"], o["tick"], o["results"], o["state_hash"][0])
    ext = [run_case(lambda cfg: OraclePool(cfg, threads=1), fns, c) for c in CASES_EXT]
    with open(os.path.join(HERE, "scenarios_ext.json"), "w") as f:
        json.dump(ext, f, indent=1)
    for o in ext:
        print(o["name"], o["tick"], o["results"], o["state_hash"][0])
