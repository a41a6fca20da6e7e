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


CGO_SRC = os.path.join(ROOT, "tests", "facade", "cgo_shape.c")


def build_c(libdir, libname, out):
    """gcc, not g++: cgo compiles its preamble as C, so include/gsim.h must be a C header."""
    cmd = ["gcc", "-O1", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", out, CGO_SRC,
           "-L" + libdir, "-l" + libname, "-lpthread", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return out


def run_c(binary):
    r = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS cgo call shape" in r.stdout, r.stdout


def test_cgo_call_shape():
    """The ABI driven the way third_party/gsim-go's cgo binding drives it (plain C, caller-allocated
    buffers with sizing calls, inputs destroyed after each call, 8 concurrent callers on one pool)."""
    run_c(build_c(os.path.join(ROOT, "tests", "hostemu"), "gsim_hostemu",
                  os.path.join(ROOT, "tests", "hostemu", "cgo_shape_check")))


@pytest.mark.gpu
def test_cgo_call_shape_on_cuda():
    run_c(build_c(os.path.join(ROOT, "consul_b200"), "gsim", os.path.join(ROOT, "tests", "facade", "cgo_shape_cuda")))


def test_go_facade_mirrors_the_identifiers_consul_uses():
    """third_party/gsim-go cannot be compiled here (no Go toolchain), but it must at least declare every
    serf / memberlist / coordinate identifier SURVEY §1 and §8(b) list, and its cgo calls must name real
    functions of include/gsim.h."""
    import re
    base = os.path.join(ROOT, "third_party", "gsim-go")
    def src(*parts):
        d = os.path.join(base, *parts)
        return "\n".join(open(os.path.join(d, f)).read() for f in sorted(os.listdir(d)) if f.endswith(".go"))
    serf, ml, co = src("serf", "serf"), src("memberlist"), src("serf", "coordinate")
    for ident in ("func Create(", "func DefaultConfig(", "func (c *Config) Init(", "func (s *Serf) Join(", "func (s *Serf) Leave(",
                  "func (s *Serf) Shutdown(", "func (s *Serf) ShutdownCh(", "func (s *Serf) UserEvent(", "func (s *Serf) Members(",
                  "func (s *Serf) LocalMember(", "func (s *Serf) NumNodes(", "func (s *Serf) SetTags(", "func (s *Serf) Stats(",
                  "func (s *Serf) RemoveFailedNode(", "func (s *Serf) RemoveFailedNodePrune(", "func (s *Serf) GetCoordinate(",
                  "func (s *Serf) GetCachedCoordinate(", "func (s *Serf) KeyManager(", "type MergeDelegate interface",
                  "type ReconnectTimeoutOverrider interface", "type MemberEvent struct", "type UserEvent struct",
                  "StatusNone MemberStatus = iota", "EventMemberReap", "EventQuery", "ListKeysWithOptions", "KeyRequestOptions",
                  "ReconnectTimeoutOverride ", "EnableNameConflictResolution", "RejoinAfterLeave", "SnapshotPath"):
        assert ident in serf, ident
    for ident in ("type Config struct", "func DefaultLANConfig(", "func DefaultWANConfig(", "type NodeAwareTransport interface",
                  "type Address struct", "func NewNetTransport(", "type NetTransportConfig struct", "func NewKeyring(",
                  "func ValidateKey(", "func ParseCIDRs(", "func LogConn(", "DisableTcpPingsForNode", "RequireNodeNames",
                  "CIDRsAllowed", "DeadNodeReclaimTime", "GossipVerifyIncoming", "func (k *Keyring) GetPrimaryKey("):
        assert ident in ml, ident
    for ident in ("type Coordinate struct", "func NewCoordinate(", "func DefaultConfig(", "func (c *Coordinate) DistanceTo("):
        assert ident in co, ident
    header = open(os.path.join(ROOT, "include", "gsim.h")).read()
    called = set(re.findall(r"C\.(gsim_[a-z0-9_]+)\(", serf))
    assert len(called) >= 15
    for fn in called:
        assert re.search(r"\b" + fn + r"\s*\(", header), fn + " is not declared in include/gsim.h"
