"""Quiet windows on the GPU (bodies in tests/test_windows_cpu.py): the window kernel — per-probe loop,
closed form on a pristine pool, generic tail — against one launch per tick and against the oracle."""
import pytest

import test_windows_cpu as wc
from consul_b200.pool import consul_test_config, lan_config

pytestmark = pytest.mark.gpu


def test_steady_state_runs_in_windows(cuda_lib):
    wc.test_steady_state_runs_in_windows(cuda_lib)


def test_join_cascade_then_windows(cuda_lib):
    wc.test_join_cascade_then_windows(cuda_lib)


@pytest.mark.parametrize("cfg_fn,ppm,ticks", [(lan_config, 20000, 700), (consul_test_config, 100000, 200)])
def test_crashes_bound_the_windows(cuda_lib, cfg_fn, ppm, ticks):
    wc.test_crashes_bound_the_windows(cuda_lib, cfg_fn, ppm, ticks)


def test_pristine_pool_closed_form_ring_passes(cuda_lib):
    """1000 members: a ring pass is 10 000 ticks, so 22 000 ticks cross two pass ends and every member's own entry"""
    wc.test_pristine_pool_closed_form(cuda_lib, 1000, 32)


def test_pristine_pool_closed_form_200k(cuda_lib):
    wc.test_pristine_pool_closed_form(cuda_lib, 200_000, 33, chunks=(100, 2560, 7, 5000, 1, 2559))


def test_closed_form_stops_at_a_member_somebody_has_not_heard_of(cuda_lib):
    wc.test_closed_form_stops_at_a_member_somebody_has_not_heard_of(cuda_lib)
