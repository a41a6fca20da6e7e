"""Randomised operation sequences on the GPU against the oracle (tests/fuzz_ops.py)."""
import pytest

import fuzz_ops
from consul_b200.pool import Pool
from oracle_binding import OraclePool

pytestmark = pytest.mark.gpu


@pytest.fixture()
def make(cuda_lib):
    return lambda cfg: [Pool(cfg, cuda_lib), OraclePool(cfg, threads=1)]


@pytest.mark.parametrize("block", range(3))
def test_random_sequences(make, cuda_lib, block):
    for seed in list(range(block * 25, block * 25 + 25)) + ([251] if block == 0 else []):
        fuzz_ops.run_sequence(make, cuda_lib, seed)


def test_calm_sequences_spend_their_ticks_in_quiet_windows(make, cuda_lib):
    """(tests/test_fuzz_parity_cpu.py) the window kernel, closed form included, under random operations"""
    for seed in range(3000, 3008):
        fuzz_ops.run_sequence(make, cuda_lib, seed, n_ops=40, calm=True)
