"""Periodic push-pull anti-entropy (GSIM_FLAG_PUSH_PULL) on the GPU against the oracle; scenario
bodies shared with the CPU suite (tests/test_pushpull_cpu.py).  Sorted after the other GPU files:
the feature is opt-in and was developed after round 1's GPU budget had been spent."""
import pytest

import test_pushpull_cpu as pp
from consul_b200.pool import FLAG_PUSH_PULL, Pool
from oracle_binding import OraclePool

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make(cuda_lib):
    return lambda cfg: [Pool(cfg, cuda_lib), OraclePool(cfg, threads=0)]


def test_push_pull_completes_what_gossip_strands(make, cuda_lib):
    pp.test_push_pull_completes_what_gossip_strands(make, cuda_lib)


def test_push_pull_period_and_stat(make, cuda_lib):
    pp.test_push_pull_period_and_stat(make, cuda_lib)


def test_push_pull_small_cluster_per_member_phases(make, cuda_lib):
    pp.test_push_pull_small_cluster_per_member_phases(make, cuda_lib)


def test_push_pull_with_wan_latency(make, cuda_lib):
    pp.test_push_pull_with_wan_latency(make, cuda_lib)


def test_snapshot_restore_mid_exchange(cuda_lib):
    pp.test_snapshot_restore_mid_exchange(cuda_lib)


def test_push_pull_100k(make, cuda_lib):
    pp.stranded_event(make, cuda_lib, FLAG_PUSH_PULL, n=100_000, ticks=600, every=100)
