"""Golden fixtures (tests/golden/scenarios.json, produced by tests/golden/make_golden.py from
the oracle): convergence ticks, stats and the 256-bit state digest of eight scripted scenarios.
CPU: the oracle still reproduces them (semantic regression pin) and so does the kernel body
compiled for the host.  GPU: the CUDA path through the C ABI reproduces them bit for bit."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402

with open(os.path.join(HERE, "golden", "scenarios.json")) as f:
    GOLDEN = {g["name"]: g for g in json.load(f)}
with open(os.path.join(HERE, "golden", "scenarios_ext.json")) as f:
    GOLDEN.update({g["name"]: g for g in json.load(f)})
ALL_CASES = mg.CASES + mg.CASES_EXT


def check(make_pool, lib, case):
    got = mg.run_case(make_pool, mg.config_fns(lib), case)
    want = GOLDEN[case[0]]
    assert got["tick"] == want["tick"]
    assert got["results"] == want["results"]
    diffs = {k: (got["stats"][k], want["stats"][k]) for k in want["stats"] if got["stats"][k] != want["stats"][k]}
    assert not diffs, diffs
    assert got["state_hash"] == want["state_hash"]


@pytest.mark.parametrize("case", ALL_CASES, ids=[c[0] for c in ALL_CASES])
def test_oracle_reproduces_golden(case):
    from consul_b200 import _lib
    from oracle_binding import OraclePool
    check(lambda cfg: OraclePool(cfg, threads=2), _lib.lib(), case)


@pytest.mark.parametrize("case", ALL_CASES, ids=[c[0] for c in ALL_CASES])
def test_kernel_body_on_host_reproduces_golden(case, hostemu_lib):
    from consul_b200.pool import Pool
    check(lambda cfg: Pool(cfg, hostemu_lib), hostemu_lib, case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", mg.CASES, ids=[c[0] for c in mg.CASES])
def test_cuda_reproduces_golden(case, cuda_lib):
    from consul_b200.pool import Pool
    check(lambda cfg: Pool(cfg, cuda_lib), cuda_lib, case)
