"""Consul calls serf from many goroutines (event loops, the Flood ticker, the leader
loop, RPC and HTTP handlers), so libgsim serialises every entry point on one mutex per pool.  Eight
host threads hammer ONE pool through the C ABI — reads, events, joins and steps interleaved — and the
pool must stay consistent: no error codes, no lost operation, invariants hold, and a second pool
that replays the same operations in their actual (serialised) order reaches the same digest."""
import threading

import pytest

from consul_b200.pool import Pool, lan_config
from parity import check_invariants

pytestmark = pytest.mark.gpu


def hammer(pool, n, n_threads=8, rounds=8):      # 8 x (alive + join intent + event) = 24 of the 30 broadcast slots
    errors, log, order = [], [], threading.Lock()

    def worker(k):
        try:
            for r in range(rounds):
                if k == 0:
                    with order:                       # the log must record the order the pool saw
                        pool.step(3)
                        log.append(("step", 3))
                elif k == 1:
                    with order:
                        x = pool.member_add()
                        ok = pool.join(x, [0])
                        log.append(("join", x, ok))
                elif k == 2:
                    with order:
                        s = pool.user_event(5 + r, b"ev%d" % r, b"p", False)
                        log.append(("event", 5 + r, b"ev%d" % r, s))
                else:                                 # readers never take the test's lock: they race the writers
                    m = pool.members(k)
                    assert len(m) >= n
                    st = pool.stats()
                    assert st["n_members"] >= n and st["tick"] <= 3 * rounds
                    pool.state_hash()
                    pool.num_nodes(k)
                    pool.poll_events(16)
        except Exception as e:                        # noqa: BLE001
            errors.append((k, repr(e)))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(n_threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return errors, log


def test_eight_threads_on_one_pool(cuda_lib):
    n = 50_000
    cfg = lan_config(cuda_lib, capacity=n + 64, n_initial=n, seed=0x5EED0008)
    pool = Pool(cfg, cuda_lib)
    errors, log = hammer(pool, n)
    assert not errors, errors
    assert sum(1 for e in log if e[0] == "step") == 8 and pool.now == 24
    assert pool.stats()["n_members"] == n + 8
    check_invariants(pool, where="after 8 threads")
    # replay the writers' operations single-threaded, in the order they were serialised
    ref = Pool(cfg, cuda_lib)
    for e in log:
        if e[0] == "step":
            ref.step(e[1])
        elif e[0] == "join":
            x = ref.member_add()
            assert x == e[1] and ref.join(x, [0]) == e[2]
        else:
            assert ref.user_event(e[1], e[2], b"p", False) == e[3]
    assert ref.state_hash() == pool.state_hash() and ref.stats() == pool.stats()
