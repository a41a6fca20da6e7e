"""Edge cases of the reference-facing surface on the GPU (bodies in tests/test_edge_cases_cpu.py)."""
import pytest

import test_edge_cases_cpu as ec
from consul_b200.pool import Pool
from oracle_binding import OraclePool

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make(cuda_lib):
    return lambda cfg: [Pool(cfg, cuda_lib), OraclePool(cfg, threads=0)]


def test_empty_pool_and_single_member(make, cuda_lib):
    ec.test_empty_pool_and_single_member(make, cuda_lib)


def test_thirty_one_concurrent_events(make, cuda_lib):
    ec.test_thirty_one_concurrent_events(make, cuda_lib)


def test_event_log_overflow_is_counted(make, cuda_lib):
    ec.test_event_log_overflow_is_counted(make, cuda_lib)


def test_join_that_reaches_nobody_and_repeated_operations(make, cuda_lib):
    ec.test_join_that_reaches_nobody_and_repeated_operations(make, cuda_lib)


def test_zero_length_horizons(make, cuda_lib):
    ec.test_zero_length_horizons(make, cuda_lib)
