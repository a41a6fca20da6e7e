"""Network coordinates (Vivaldi; SURVEY §8f N3, §8b GetCoordinate), opt-in with
GSIM_FLAG_COORDINATES: kernel body == oracle BIT FOR BIT on IEEE doubles (both builds disable FMA
contraction), and the embedding does what Consul uses it for — predicted round trips between
datacenters order like the real ones (agent/router/router.go:62-67 sorts by this distance)."""
import math

import numpy as np
import pytest

import scenarios as sc
from consul_b200.pool import FLAG_COORDINATES, Pool, lan_config, wan_config
from oracle_binding import OraclePool
from parity import compare_pools


@pytest.fixture()
def make(hostemu_lib):
    return lambda cfg: [Pool(cfg, hostemu_lib), OraclePool(cfg)]


def dist(a, b):
    """coordinate.DistanceTo in seconds (without the Duration round trip)."""
    va, ea, aa, ha = a
    vb, eb, ab, hb = b
    raw = math.sqrt(sum((x - y) ** 2 for x, y in zip(va, vb))) + ha + hb
    adj = raw + aa + ab
    return adj if adj > 0 else raw


def three_dc_matrix():
    # one-way latencies in ticks between 3 datacenters: 0-1 near, 0-2 far, 1-2 in between; every
    # round trip stays within WAN's ProbeTimeout (6 ticks), so every pair is measured by direct acks
    return np.array([[1, 2, 4], [2, 1, 3], [4, 3, 1]], dtype=np.uint8)


def test_bit_exact_against_oracle_and_digest(make, hostemu_lib):
    n = 3 * 128 * 2
    cfg = wan_config(hostemu_lib, capacity=n + 2, n_initial=n, seed=11, flags=FLAG_COORDINATES, mailbox_depth=8,
                     packet_loss_ppm=100000)
    pools = make(cfg)
    for p in pools:
        p.latency_set(three_dc_matrix())
    x = sc.both(pools, lambda p: p.member_add())
    sc.both(pools, lambda p: p.join(x, [0]))
    for p in pools:
        p.crash_many([5, 200])
    sc.step_compare(pools, 400, 40, "coordinates")          # the digest folds every coordinate's bits
    for i in (0, 1, 130, 300, 767, x):
        a, b = pools[0].coordinate(i), pools[1].coordinate(i)
        assert a == b, (i, a, b)                               # exact doubles
    vec, err, adj, h = pools[0].coordinate(1)
    assert any(v != 0.0 for v in vec) and 0.0 < err <= 1.5 and h >= 10.0e-6


def test_embedding_orders_datacenters_by_round_trip(hostemu_lib):
    """After a few hundred probes per member the predicted distance between members of DC0 and DC1
    (1 s round trip) is smaller than DC1-DC2 (2 s), which is smaller than DC0-DC2 (3 s), and each is
    within 35 % of the round trip Vivaldi was fed; members of one datacenter sit on top of each other."""
    n = 3 * 128 * 4
    tau = 0.5                                                 # WAN tick
    cfg = wan_config(hostemu_lib, capacity=n, n_initial=n, seed=5, flags=FLAG_COORDINATES, mailbox_depth=8)
    p = Pool(cfg, hostemu_lib)
    m = three_dc_matrix()
    p.latency_set(m)
    p.step(4000)                                              # 400 probes per member
    dc = lambda i: (i // 128) % 3
    members = {d: [i for i in range(0, n, 37) if dc(i) == d][:6] for d in range(3)}
    coords = {i: p.coordinate(i) for d in members for i in members[d]}

    def mean_pred(a, b):
        return np.mean([dist(coords[i], coords[j]) for i in members[a] for j in members[b] if i != j])

    true = lambda a, b: 0.0005 + ((m[a][b] - 1) + (m[b][a] - 1)) * tau
    d01, d12, d02, d00 = mean_pred(0, 1), mean_pred(1, 2), mean_pred(0, 2), mean_pred(0, 0)
    assert d00 < d01 < d12 < d02, (d00, d01, d12, d02)
    for pred, (a, b) in ((d01, (0, 1)), (d12, (1, 2)), (d02, (0, 2))):
        assert abs(pred - true(a, b)) / true(a, b) < 0.35, (pred, true(a, b))
    assert np.mean([coords[i][1] for i in coords]) < 0.6      # the error estimate has come down from 1.5


def test_api_contract(hostemu_lib):
    p = Pool(lan_config(hostemu_lib, capacity=300, n_initial=300, seed=1), hostemu_lib)
    with pytest.raises(Exception):
        p.coordinate(0)                                        # pool created without the flag
    q = Pool(lan_config(hostemu_lib, capacity=300, n_initial=300, seed=1, flags=FLAG_COORDINATES), hostemu_lib)
    vec, err, adj, h = q.coordinate(7)
    assert vec == [0.0] * 8 and err == 1.5 and adj == 0.0 and h == 10.0e-6   # coordinate.NewCoordinate
    with pytest.raises(Exception):
        q.coordinate(300)
    blob = q.snapshot()
    q.step(200)
    moved = q.coordinate(7)
    q.restore(blob)
    assert q.coordinate(7) == (vec, err, adj, h)
    q.step(200)
    assert q.coordinate(7) == moved                            # deterministic replay of doubles
