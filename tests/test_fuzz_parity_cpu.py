"""Randomised operation sequences (tests/fuzz_ops.py): kernel body vs oracle after every operation —
members added, joined, crashed, leaving, force-left/pruned, re-tagged, events fired, injected and
retired, reconnect overrides, with random presets, loss, UDP budgets, push-pull, reaper timers and
latency matrices.  Seed 251 is the sequence that found the stale-buffer bug in gsim_force_leave."""
import pytest

import fuzz_ops
from consul_b200.pool import Pool
from oracle_binding import OraclePool


@pytest.fixture()
def make(hostemu_lib):
    return lambda cfg: [Pool(cfg, hostemu_lib), OraclePool(cfg)]


@pytest.mark.parametrize("block", range(6))
def test_random_sequences(make, hostemu_lib, block):
    for seed in range(block * 40, block * 40 + 40):
        fuzz_ops.run_sequence(make, hostemu_lib, seed)


def test_regression_seeds(make, hostemu_lib):
    for seed in (251,):
        fuzz_ops.run_sequence(make, hostemu_lib, seed)
