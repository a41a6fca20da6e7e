"""CSR peer graph on the GPU (scenario bodies shared with tests/test_graph_cpu.py): complete CSR ==
implicit complete graph, restricted topologies against the oracle, probe ring over a row.  Also
exercises the CUDA-graph cache invalidation when the column pointers change."""
import pytest

import test_graph_cpu as tg
from consul_b200.pool import Pool
from oracle_binding import OraclePool

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make(cuda_lib):
    return lambda cfg: [Pool(cfg, cuda_lib), OraclePool(cfg, threads=0)]


def test_complete_csr_equals_implicit_complete_graph(cuda_lib):
    tg.test_complete_csr_equals_implicit_complete_graph("kernel-body", cuda_lib)


def test_segment_ring_parity_and_reachability(make, cuda_lib):
    tg.test_segment_ring_parity_and_reachability(make, cuda_lib)


def test_disconnected_components_do_not_leak(make, cuda_lib):
    tg.test_disconnected_components_do_not_leak(make, cuda_lib)


def test_graph_api_contract(cuda_lib):
    tg.test_graph_api_contract(cuda_lib)


def test_graph_attached_after_cuda_graphs_were_captured(make, cuda_lib):
    """Run 200 ticks on the complete graph first (64-tick CUDA graphs get captured), then attach a
    CSR: the cached launches must not keep the old column pointers."""
    import numpy as np
    import scenarios as sc
    from consul_b200.pool import lan_config
    n = 1024
    pools = make(lan_config(cuda_lib, capacity=n, n_initial=n, seed=17))
    sc.step_compare(pools, 200, 100, "before")
    rp, ci = tg.segment_ring_csr(n, 128)
    for p in pools:
        p.graph_set(rp, ci)
    sc.both(pools, lambda p: p.user_event(0, b"e", b"p", False))
    sc.step_compare(pools, 256, 64, "after attach")
    for p in pools:
        p.graph_set(None, None)
    sc.step_compare(pools, 128, 64, "after detach")
