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


def test_torch_free_quickcheck_binary():
    """tests/facade/gpu_quickcheck.cpp: libgsim.so against liboracle.so, no Python in the loop."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "tests", "facade", "gpu_quickcheck_cuda")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "tests", "facade", "gpu_quickcheck.cpp"),
                    "-L" + os.path.join(root, "consul_b200"), "-lgsim", "-L" + os.path.join(root, "oracle"), "-loracle",
                    "-Wl,-rpath," + os.path.join(root, "consul_b200"), "-Wl,-rpath," + os.path.join(root, "oracle"),
                    "-o", out], check=True)
    r = subprocess.run([out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout + r.stderr


def test_invariants_at_scale(cuda_lib):
    """4 Mi members — beyond what the oracle is asked to follow tick by tick: a join, two events and
    a 5 % crash wave with 10 % packet loss; the state invariants (tests/parity.py check_invariants:
    queued within heard, transmits within the limit, counters equal to recounts, no probe action in
    the past, incarnations and Lamport clocks monotone) hold at every checkpoint."""
    from consul_b200.pool import lan_config
    from parity import check_invariants
    n = 4 * 1024 * 1024
    p = Pool(lan_config(cuda_lib, capacity=n + 2, n_initial=n, seed=0x5EED0077, packet_loss_ppm=100000), cuda_lib)
    x = p.member_add()
    p.join(x, [0])
    p.user_event(5, b"deploy", b"x" * 32, False)
    snap = check_invariants(p, None, "t=0")
    for k, ticks in enumerate((7, 64, 129, 300)):
        if k == 1:
            assert p.crash_fraction(50000, 1) > 0
            p.user_event(9, b"second", b"", False)
        p.step(ticks)
        snap = check_invariants(p, snap, f"checkpoint {k}")
    s = p.stats()
    assert s["suspects"] > 0 and s["deads"] > 0 and s["rumors_accepted"] > 2 * n


def test_snapshot_resume_is_bit_exact(cuda_lib):
    """Checkpoint / resume through the plane-compressed snapshot (device fills for planes of one
    repeated word) reproduces an uninterrupted run bit for bit."""
    import test_gpu_parity as gp
    assert gp.determinism_run(cuda_lib) == gp.determinism_run(cuda_lib, restore_at=100)
