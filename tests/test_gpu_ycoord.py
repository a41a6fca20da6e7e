"""Network coordinates on the GPU: IEEE doubles bit-identical to the oracle (FMA contraction is off
in both builds), embedding quality, API contract — bodies in tests/test_coordinates_cpu.py."""
import pytest

import test_coordinates_cpu as tc
from consul_b200.pool import Pool
from oracle_binding import OraclePool

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make(cuda_lib):
    return lambda cfg: [Pool(cfg, cuda_lib), OraclePool(cfg, threads=0)]


def test_bit_exact_against_oracle_and_digest(make, cuda_lib):
    tc.test_bit_exact_against_oracle_and_digest(make, cuda_lib)


def test_embedding_orders_datacenters_by_round_trip(cuda_lib):
    tc.test_embedding_orders_datacenters_by_round_trip(cuda_lib)


def test_api_contract(cuda_lib):
    tc.test_api_contract(cuda_lib)
