"""The reference's gossip tests replayed through the C++ serf facade (include/gsim_serf.hpp):
tests/facade/facade_check.cpp.  CPU: linked with the host emulation; GPU: with libgsim.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "facade", "facade_check.cpp")


def build(libdir, libname, out):
    cmd = ["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", out, SRC,
           "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return out


def run(binary, extended=False):
    r = subprocess.run([binary] + (["--extended"] if extended else []), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS" in r.stdout
    names = ("TestServer_LANReap (reaper)", "TestServer_JoinWAN", "TestServer_WANReap", "Serf.SetTags", "TestMerge_LAN",
             "TestClient_ShortReconnectTimeout", "GetCoordinate") if extended else \
        ("TestServer_JoinLAN", "TestServer_LANReap", "TestClientServer_UserEvent", "TestAgent_Leave")
    for name in names:
        assert "PASS " + name in r.stdout


def test_facade_on_host_emulation():
    binary = build(os.path.join(ROOT, "tests", "hostemu"), "gsim_hostemu",
                   os.path.join(ROOT, "tests", "hostemu", "serf_facade_check"))
    run(binary)
    run(binary, extended=True)


@pytest.mark.gpu
def test_facade_on_cuda():
    run(build(os.path.join(ROOT, "consul_b200"), "gsim",
              os.path.join(ROOT, "tests", "facade", "facade_check_cuda")))
