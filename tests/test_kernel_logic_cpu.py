"""CPU-side parity: the tick kernel's row body (compiled by g++, tests/hostemu) against the
oracle, field by field after every phase, including adversarial row orders.  This is the
pre-GPU gate; the parity claim proper is tests/test_gpu_parity.py on a B200."""
import os
import subprocess
import sys

import pytest

import scenarios as sc
from consul_b200.pool import Pool
from oracle_binding import OraclePool


@pytest.fixture()
def make(hostemu_lib):
    return lambda cfg: [Pool(cfg, hostemu_lib), OraclePool(cfg)]


def test_c1_three_node_join(make, hostemu_lib):
    for seed in (1, 2, 3):
        t, td = sc.c1_three_node_join(make, hostemu_lib, seed)
        assert 0 < t < 40 and td > 0


@pytest.mark.parametrize("n", [1, 2, 3, 17, 1000, 20000])
def test_c2_join_cascade(make, hostemu_lib, n):
    t = sc.c2_join_cascade(make, hostemu_lib, n, every=4)
    assert t < 80


def test_c3_crash_lan(make, hostemu_lib):
    crashed, t = sc.c3_crash(make, hostemu_lib, 3000)
    assert 250 < crashed < 350


def test_c3_crash_fast_timers(make, hostemu_lib):
    # consul test timing: k = 0, short suspicion
    sc.c3_crash(make, hostemu_lib, 500, ppm=200000, cfg_fn=sc.consul_test_config, every=5)
    sc.c3_crash(make, hostemu_lib, 400, ppm=500000, cfg_fn=sc.wan_config, every=40)


@pytest.mark.parametrize("n", [2, 50, 5000])
def test_c4_user_event(make, hostemu_lib, n):
    sc.c4_user_event(make, hostemu_lib, n)


def test_leave(make, hostemu_lib):
    sc.leave_scenario(make, hostemu_lib)


def test_lossy(make, hostemu_lib):
    s = sc.lossy_scenario(make, hostemu_lib)
    # the TCP fallback ping saves every live target: loss alone never causes a false suspicion
    assert s["refutes"] == 0 and s["nacks"] > 0 and s["suspects"] == 3
    s = sc.lossy_scenario(make, hostemu_lib, n=200, loss_ppm=400000, seed=12, ticks=400, disable_tcp_pings=1)
    assert s["refutes"] > 0 and s["confirmations"] > 0


def test_udp_budget_ordering(make, hostemu_lib):
    sc.budget_scenario(make, hostemu_lib)


def test_event_window(make, hostemu_lib):
    sc.event_window_scenario(make, hostemu_lib)


@pytest.mark.parametrize("order", ["1", "2"])
def test_row_order_independence(order):
    """Rows executed in reverse / odd-even order must give identical results (the property
    that makes the CUDA launch deterministic under any block scheduling)."""
    code = (
        "import sys; sys.path[:0]=[%r,%r]\n"
        "import scenarios as sc\n"
        "from consul_b200 import _lib\n"
        "from consul_b200.pool import Pool\n"
        "from oracle_binding import OraclePool\n"
        "L=_lib.load(%r)\n"
        "mk=lambda cfg:[Pool(cfg,L),OraclePool(cfg)]\n"
        "sc.lossy_scenario(mk,L,n=300,ticks=200)\n"
        "sc.c3_crash(mk,L,800,ppm=150000)\n"
        "sc.c2_join_cascade(mk,L,2000)\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
         os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu", "libgsim_hostemu.so"))
    env = dict(os.environ, GSIM_HOSTEMU_ORDER=order)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_oracle_threads_agree():
    """The OpenMP oracle (used as the CPU baseline) is bit-identical to the 1-thread oracle."""
    from consul_b200 import _lib
    L = _lib.lib()
    mk = lambda cfg: [OraclePool(cfg, threads=1), OraclePool(cfg, threads=4)]
    sc.lossy_scenario(mk, L, n=400, ticks=200)
    sc.c3_crash(mk, L, 1500)


def test_lan_reap(make, hostemu_lib):
    """serf's reaper with the timings of TestServer_LANReap."""
    for seed in (1, 2):
        sc.lan_reap_scenario(make, hostemu_lib, seed)


def test_set_tags_update(make, hostemu_lib):
    """serf.SetTags / memberlist.UpdateNode / EventMemberUpdate."""
    sc.set_tags_scenario(make, hostemu_lib)


def test_quickcheck_harness_on_host_emulation(hostemu_lib, tmp_path):
    """The torch-free C++ parity harness (tests/facade/gpu_quickcheck.cpp) itself, linked against the
    host emulation: all seven scenarios agree with the oracle."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "quickcheck_cpu")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "tests", "facade", "gpu_quickcheck.cpp"),
                    "-L" + os.path.join(root, "tests", "hostemu"), "-lgsim_hostemu", "-L" + os.path.join(root, "oracle"),
                    "-loracle", "-Wl,-rpath," + os.path.join(root, "tests", "hostemu"),
                    "-Wl,-rpath," + os.path.join(root, "oracle"), "-o", out], check=True)
    r = subprocess.run([out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("PASS ") == 7 and "ALL PASS" in r.stdout, r.stdout + r.stderr


def test_short_reconnect_timeout(make, hostemu_lib):
    """serf.Config.ReconnectTimeoutOverride with the timings of TestClient_ShortReconnectTimeout."""
    for seed in (1, 2):
        sc.short_reconnect_timeout_scenario(make, hostemu_lib, seed)
