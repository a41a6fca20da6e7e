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


@pytest.mark.parametrize("block", range(2))
def test_calm_sequences_spend_their_ticks_in_quiet_windows(make, hostemu_lib, block):
    """No loss, mostly time passing in steps of up to 6000 ticks on pools of 3 .. 700 members: quiet windows,
    the closed form of pristine pools, ring passes ending, own ring entries, joiners that stay pending."""
    for seed in range(3000 + block * 30, 3030 + block * 30):
        fuzz_ops.run_sequence(make, hostemu_lib, seed, n_ops=40, calm=True)
