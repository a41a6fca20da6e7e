"""BASELINE config 5 on the GPU: WAN latency pools (deeper mailbox ring, latency matrix, slow acks)
and the two-pool federation with bridge members — the sm_100a path through the C ABI against the
oracle, bit for bit, then the full-size run (2 x 8 Mi members) through size-independent
properties.  The scenario bodies are the ones the CPU suite runs on the host emulation
(tests/test_latency_cpu.py)."""
import numpy as np
import pytest

import test_latency_cpu as wl
from consul_b200.pool import Pool, lan_config, wan_config
from consul_b200.wan import WanFederation
from oracle_binding import OraclePool
from parity import compare_pools

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make(cuda_lib):
    return lambda cfg: [Pool(cfg, cuda_lib), OraclePool(cfg, threads=0)]


def test_all_ones_matrix_is_the_plain_model(cuda_lib):
    wl.test_all_ones_matrix_is_the_plain_model(cuda_lib)


@pytest.mark.parametrize("cfg_fn,n_dcs", [(wan_config, 64), (lan_config, 5)])
def test_event_dissemination_parity(make, cuda_lib, cfg_fn, n_dcs):
    wl.test_event_dissemination_parity(make, cuda_lib, cfg_fn, n_dcs)


def test_slow_acks_use_the_indirect_stage(make, cuda_lib):
    wl.test_slow_acks_use_the_indirect_stage(make, cuda_lib)


def test_lossy_crash_parity_with_latency(make, cuda_lib):
    wl.test_lossy_crash_parity_with_latency(make, cuda_lib)


def test_snapshot_restore_with_packets_in_flight(cuda_lib):
    wl.test_snapshot_restore_with_packets_in_flight(cuda_lib)


def test_c5_two_pool_federation_parity(cuda_lib):
    wl.test_c5_two_pool_federation_parity(cuda_lib)


def test_c5_full_size_properties(cuda_lib):
    """2 pools x 8 388 608 members, 64 datacenters, 5 bridges each, the C5 matrix, WAN timing:
    the event fired at A/DC0 is delivered exactly once by every member of both pools, every
    event clock has witnessed it, nobody is suspected, and B was seeded by bridges only."""
    n = 8 * 1024 * 1024
    cfg = lambda seed: wan_config(cuda_lib, capacity=n, n_initial=n, seed=seed, mailbox_depth=8)
    fed = WanFederation(Pool(cfg(0x5EED0051), cuda_lib), Pool(cfg(0x5EED0052), cuda_lib), n_dcs=64,
                        bridges_per_dc=5, n_members=n)
    assert len(fed.bridges) == 320
    fed.fire(0, 7, b"deploy", b"x" * 32)                     # member 7: DC0, not a bridge
    t = fed.run_until_converged(b"deploy", b"x" * 32, 600)
    assert t is not None and 10 < t < 200
    key = (b"deploy", b"x" * 32)
    for x, p in enumerate(fed.pools):
        info = p.rumor_info(fed.slots[key][x])
        assert info["heard_count"] == n
        s = p.stats()
        # exactly-once delivery: the origin, the bridge re-fires, and gossip add up to n
        assert s["rumors_accepted"] == n - (1 if x == 0 else 0) - fed.forwarded_into[x]
        assert s["suspects"] == 0 and s["refutes"] == 0 and s["probe_failures"] == 0
        assert int(p.column("ltime_event")[:n].min()) >= 2
    assert fed.pools[1].rumor_info(fed.slots[key][1])["origin"] in fed.bridges
    assert fed.forwarded >= 1
    # draining the queues changes nothing that was delivered
    fed.step(8)
    for x, p in enumerate(fed.pools):
        assert p.rumor_info(fed.slots[key][x])["heard_count"] == n



def test_lan_reap(make, cuda_lib):
    """SURVEY 8a row a17 on the GPU: TestServer_LANReap timings, reaper pass as a device kernel."""
    import scenarios as sc
    for seed in (1, 2):
        sc.lan_reap_scenario(make, cuda_lib, seed)


def test_facade_extended_on_cuda():
    """TestServer_LANReap with the reaper's own timers and TestServer_JoinWAN through the C++ serf
    facade on libgsim.so (the round-1 facade set runs in tests/test_facade.py)."""
    import os
    import test_facade as tf
    binary = tf.build(os.path.join(tf.ROOT, "consul_b200"), "gsim",
                      os.path.join(tf.ROOT, "tests", "facade", "facade_check_cuda"))
    tf.run(binary, extended=True)


def test_set_tags_update(make, cuda_lib):
    """(*Serf).SetTags -> EventMemberUpdate on the GPU against the oracle."""
    import scenarios as sc
    sc.set_tags_scenario(make, cuda_lib)


@pytest.mark.parametrize("case_name", ["wan_c5_latency_event_16k", "wan_slow_links_lossy_3k", "pushpull_stranded_3k",
                                       "lan_reap_3", "set_tags_500"])
def test_cuda_reproduces_extended_golden(case_name, cuda_lib):
    """tests/golden/scenarios_ext.json (oracle-generated): ticks, counters and the 256-bit digest."""
    import test_golden as tg
    case = [c for c in tg.mg.CASES_EXT if c[0] == case_name][0]
    tg.check(lambda cfg: Pool(cfg, cuda_lib), cuda_lib, case)


def test_short_reconnect_timeout(make, cuda_lib):
    """serf.Config.ReconnectTimeoutOverride (TestClient_ShortReconnectTimeout timings) on the GPU."""
    import scenarios as sc
    sc.short_reconnect_timeout_scenario(make, cuda_lib, 1)
