// facade_check.cpp — the reference's own gossip tests, replayed against the C++ serf facade
// (include/gsim_serf.hpp) over the C ABI.  Linked twice by tests/test_facade.py: against the
// host emulation (CPU suite) and against libgsim.so (GPU suite).
//
//   TestServer_JoinLAN            agent/consul/server_test.go:509-529
//   TestServer_LANReap (shape)    agent/consul/server_test.go:666-733
//   TestAgent_ForceLeave          agent/agent_endpoint_test.go:2524-2566
//   TestClientServer_UserEvent    agent/consul/client_test.go:756-835
//   TestAgent_Leave               agent/agent_endpoint_test.go:2447
#include <cstdio>
#include <cstdlib>
#include <set>
#include <string>

#include "../../include/gsim_serf.hpp"

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                        \
    }                                                                      \
  } while (0)

using namespace serf;

static int count_status(Serf& s, int status) {
  int n = 0;
  for (auto& m : s.Members()) n += m.Status == status;
  return n;
}

// retry.Run (sdk/testutil/retry/retryer.go:18-20: 7 s / 25 ms) in simulated time
template <class F>
static bool eventually(Pool& p, uint32_t max_ticks, F f) {
  for (uint32_t t = 0; t < max_ticks; ++t) {
    if (f()) return true;
    p.Step(1);
  }
  return f();
}

static gsim_config test_cfg() {
  gsim_config c = Pool::TestConfig();
  c.capacity = 16;
  c.n_initial = 0;
  c.seed = 42;
  c.flags = GSIM_FLAG_LOG_GLOBAL_EVENTS;
  return c;
}

static void test_join_lan() {
  Pool pool(test_cfg());
  Config c1, c2;
  c1.NodeName = "s1";
  c1.Tags = {{"role", "consul"}, {"dc", "dc1"}};
  c2.NodeName = "s2";
  c2.Tags = {{"role", "consul"}, {"dc", "dc1"}};
  auto s1 = Serf::Create(pool, c1);
  auto s2 = Serf::Create(pool, c2);
  CHECK(s1->Members().size() == 1 && s2->Members().size() == 1);
  CHECK(s2->Join({"s1/127.0.0.1:8301"}, true) == 1);
  CHECK(eventually(pool, 140, [&] { return s1->Members().size() == 2 && s2->Members().size() == 2; }));
  CHECK(s1->LocalMember().Name == "s1" && s1->LocalMember().Status == StatusAlive);
  CHECK(s1->Members()[1].Tags.at("role") == "consul");
  bool threw = false;
  try {
    s2->Join({"nosuch/127.0.0.1:1"}, true);
  } catch (const Error&) {
    threw = true;
  }
  CHECK(threw);
  threw = false;
  try {
    Serf::Create(pool, c1);  // node name conflict
  } catch (const Error&) {
    threw = true;
  }
  CHECK(threw);
  std::puts("PASS TestServer_JoinLAN");
}

static void test_lan_reap_and_force_leave() {
  Pool pool(test_cfg());
  std::deque<Event> ch1;
  Config c1, c2, c3;
  c1.NodeName = "s1";
  c1.EventCh = &ch1;
  c2.NodeName = "s2";
  c3.NodeName = "s3";
  auto s1 = Serf::Create(pool, c1), s2 = Serf::Create(pool, c2), s3 = Serf::Create(pool, c3);
  s2->Join({"s1/x"}, true);
  s3->Join({"s1/x"}, true);
  CHECK(eventually(pool, 140, [&] { return s1->NumNodes() == 3 && s2->NumNodes() == 3 && s3->NumNodes() == 3; }));
  s3->Shutdown();  // crash without Leave
  CHECK(eventually(pool, 400, [&] { return count_status(*s1, StatusFailed) == 1 && count_status(*s2, StatusFailed) == 1; }));
  CHECK(count_status(*s1, StatusAlive) == 2);
  pool.PumpEvents();
  std::set<std::string> joined;
  bool failed_seen = false;
  for (auto& e : ch1) {
    if (e.Type == EventMemberJoin) joined.insert(e.Members[0].Name);
    if (e.Type == EventMemberFailed && e.Members[0].Name == "s3") failed_seen = true;
  }
  CHECK(joined.count("s2") && joined.count("s3") && failed_seen);
  s1->RemoveFailedNode("s3");  // Failed -> Left
  CHECK(count_status(*s2, StatusLeft) == 1 && count_status(*s2, StatusFailed) == 0);
  s1->RemoveFailedNodePrune("s3");  // erased (EventMemberReap in serf)
  CHECK(s1->Members().size() == 2);
  std::puts("PASS TestServer_LANReap / TestAgent_ForceLeave");
}

// TestServer_LANReap with its own timings (server_test.go:675-677): nobody calls RemoveFailedNode;
// the reaper forgets s2 once it has been Failed for ReconnectTimeout.
static void test_lan_reap_timers() {
  gsim_config c = test_cfg();
  c.reconnect_timeout_ns = 250ull * 1000000;
  c.tombstone_timeout_ns = 250ull * 1000000;
  c.reap_interval_ns = 300ull * 1000000;
  Pool pool(c);
  std::deque<Event> ch1;
  Config c1, c2, c3;
  c1.NodeName = "s1";
  c1.EventCh = &ch1;
  c2.NodeName = "s2";
  c3.NodeName = "s3";
  auto s1 = Serf::Create(pool, c1), s2 = Serf::Create(pool, c2), s3 = Serf::Create(pool, c3);
  s2->Join({"s1/x"}, true);
  s3->Join({"s1/x"}, true);
  CHECK(eventually(pool, 140, [&] { return s1->Members().size() == 3 && s2->Members().size() == 3 && s3->Members().size() == 3; }));
  s2->Shutdown();
  CHECK(eventually(pool, 400, [&] { return s1->Members().size() == 2 && s3->Members().size() == 2; }));
  CHECK(count_status(*s1, StatusAlive) == 2 && count_status(*s1, StatusFailed) == 0);
  pool.PumpEvents();
  int failed_at = -1, reaped_at = -1, k = 0;
  for (auto& e : ch1) {
    if (e.Type == EventMemberFailed && e.Members[0].Name == "s2") failed_at = k;
    if (e.Type == EventMemberReap && e.Members[0].Name == "s2") reaped_at = k;
    ++k;
  }
  CHECK(failed_at >= 0 && reaped_at > failed_at);
  std::puts("PASS TestServer_LANReap (reaper)");
}

// TestServer_WANReap (server_test.go:767-810): two datacenters on the WAN pool, ReconnectTimeout =
// TombstoneTimeout = 250 ms, ReapInterval = 500 ms; the second one shuts down and the first one's
// WANMembers() goes from 2 to 1 (which is what empties router.GetDatacenters()).
static void test_wan_reap() {
  gsim_config c = test_cfg();
  c.reconnect_timeout_ns = 250ull * 1000000;
  c.tombstone_timeout_ns = 250ull * 1000000;
  c.reap_interval_ns = 500ull * 1000000;
  Pool pool(c);
  Config c1, c2;
  c1.NodeName = "s1.dc1";
  c1.Tags = {{"role", "consul"}, {"dc", "dc1"}};
  c2.NodeName = "s2.dc2";
  c2.Tags = {{"role", "consul"}, {"dc", "dc2"}};
  auto s1 = Serf::Create(pool, c1), s2 = Serf::Create(pool, c2);
  CHECK(s2->Join({"s1.dc1/127.0.0.1:8302"}, true) == 1);
  CHECK(eventually(pool, 140, [&] { return s1->Members().size() == 2 && s2->Members().size() == 2; }));
  s2->Shutdown();
  CHECK(eventually(pool, 400, [&] { return s1->Members().size() == 1; }));
  CHECK(s1->Members()[0].Name == "s1.dc1" && s1->Members()[0].Tags.at("dc") == "dc1");
  std::puts("PASS TestServer_WANReap");
}

// TestServer_JoinWAN (server_test.go:735-810): one server per datacenter, joined over the WAN
// pool (names carry the datacenter, server_serf.go:90-93) with memberlist's WAN timing.
static void test_join_wan() {
  gsim_config c = Pool::DefaultWANConfig();
  c.capacity = 16;
  c.n_initial = 0;
  c.seed = 43;
  c.phase_group = 1;
  Pool pool(c);
  Config c1, c2;
  c1.NodeName = "s1.dc1";
  c1.Tags = {{"role", "consul"}, {"dc", "dc1"}};
  c2.NodeName = "s2.dc2";
  c2.Tags = {{"role", "consul"}, {"dc", "dc2"}};
  auto s1 = Serf::Create(pool, c1), s2 = Serf::Create(pool, c2);
  CHECK(s2->Join({"s1.dc1/127.0.0.1:8302"}, true) == 1);
  CHECK(eventually(pool, 100, [&] { return s1->Members().size() == 2 && s2->Members().size() == 2; }));
  std::set<std::string> dcs;
  for (auto& m : s1->Members()) dcs.insert(m.Tags.at("dc"));
  CHECK(dcs.count("dc1") && dcs.count("dc2"));
  CHECK(s1->Stats().at("members") == "2");
  std::puts("PASS TestServer_JoinWAN");
}

// (*Serf).SetTags (libserf/serf.go:51; agent/consul/server_serf.go:101-146 builds the tag map):
// the other members see the new tags and get exactly one EventMemberUpdate.
static void test_set_tags() {
  Pool pool(test_cfg());
  std::deque<Event> ch1, ch3;
  Config c1, c2, c3;
  c1.NodeName = "s1";
  c1.EventCh = &ch1;
  c2.NodeName = "s2";
  c2.Tags = {{"role", "consul"}, {"vsn", "2"}};
  c3.NodeName = "s3";
  c3.EventCh = &ch3;
  auto s1 = Serf::Create(pool, c1), s2 = Serf::Create(pool, c2), s3 = Serf::Create(pool, c3);
  s2->Join({"s1/x"}, true);
  s3->Join({"s1/x"}, true);
  CHECK(eventually(pool, 140, [&] { return s1->NumNodes() == 3 && s3->NumNodes() == 3; }));
  const uint32_t inc_before = s1->Members()[1].Incarnation;
  s2->SetTags({{"role", "consul"}, {"vsn", "3"}, {"read_replica", "1"}});
  pool.Step(60);
  pool.PumpEvents();
  int upd1 = 0, upd3 = 0;
  for (auto& e : ch1) upd1 += e.Type == EventMemberUpdate && e.Members[0].Name == "s2";
  for (auto& e : ch3) upd3 += e.Type == EventMemberUpdate && e.Members[0].Name == "s2";
  CHECK(upd1 == 1 && upd3 == 1);
  CHECK(s1->Members()[1].Tags.at("vsn") == "3" && s3->Members()[1].Tags.count("read_replica") == 1);
  CHECK(s1->Members()[1].Incarnation == inc_before + 1);
  CHECK(count_status(*s1, StatusAlive) == 3);
  std::puts("PASS Serf.SetTags / EventMemberUpdate");
}

// serf.Config.Merge: Consul's lanMergeDelegate refuses members of another datacenter
// (agent/consul/merge.go:34-88; table in merge_test.go TestMerge_LAN "client/server in the wrong
// datacenter").  The delegate is host code; what the gossip layer owes it is the call on both
// sides of the join and the cancelled merge.
static void test_merge_delegate() {
  Pool pool(test_cfg());
  auto lan_merge = [](const std::string& dc) {
    return [dc](const std::vector<Member>& members) -> std::string {
      for (auto& m : members) {
        auto it = m.Tags.find("dc");
        if (it != m.Tags.end() && it->second != dc) return "Member '" + m.Name + "' part of wrong datacenter '" + it->second + "'";
      }
      return "";
    };
  };
  Config c1, c2, c3;
  c1.NodeName = "node0";
  c1.Tags = {{"role", "consul"}, {"dc", "dc1"}};
  c1.Merge = lan_merge("dc1");
  c2.NodeName = "node1";
  c2.Tags = {{"role", "node"}, {"dc", "dc2"}};  // a client of the wrong datacenter, no delegate of its own
  c3.NodeName = "node2";
  c3.Tags = {{"role", "node"}, {"dc", "dc1"}};
  auto s1 = Serf::Create(pool, c1), s2 = Serf::Create(pool, c2), s3 = Serf::Create(pool, c3);
  bool refused = false;
  try {
    s2->Join({"node0/127.0.0.1:8301"}, true);
  } catch (const Error& e) {
    refused = std::string(e.what()).find("wrong datacenter") != std::string::npos;
  }
  CHECK(refused);
  CHECK(s3->Join({"node0/127.0.0.1:8301"}, true) == 1);  // the good cluster merges
  pool.Step(60);
  CHECK(s1->Members().size() == 2 && s3->Members().size() == 2 && s2->Members().size() == 1);
  // a mixed seed list: the refusing peer is skipped, the acceptable one is contacted
  Config c4;
  c4.NodeName = "node3";
  c4.Tags = {{"role", "node"}, {"dc", "dc2"}};
  auto s4 = Serf::Create(pool, c4);
  CHECK(s4->Join({"node0/x", "node1/x"}, true) == 1);  // node0 refuses dc2, node1 (dc2, no delegate) accepts
  pool.Step(60);
  CHECK(s4->Members().size() == 2 && s2->Members().size() == 2 && s1->Members().size() == 2);
  std::puts("PASS TestMerge_LAN (delegate called on both sides, merge cancelled)");
}

// TestClient_ShortReconnectTimeout (agent/consul/client_test.go:862-894): clients advertise
// rc_tm=100ms; libserf's ReconnectOverride (internal/gossip/libserf/serf.go:68-85) turns the tag
// into the member's reconnect timeout; ReapInterval 50 ms.
static void test_short_reconnect_timeout() {
  gsim_config c = test_cfg();
  c.reap_interval_ns = 50ull * 1000000;
  Pool pool(c);
  auto reconnect_override = [](const Member& m, uint64_t dflt) -> uint64_t {
    auto it = m.Tags.find("rc_tm");
    if (it == m.Tags.end()) return dflt;
    const uint64_t ms = std::strtoull(it->second.c_str(), nullptr, 10);  // "100ms"
    return ms ? ms * 1000000ull : dflt;
  };
  Config cs, c0, c1;
  cs.NodeName = "server";
  cs.Tags = {{"role", "consul"}};
  cs.ReconnectTimeoutOverride = reconnect_override;
  c0.NodeName = "client0";
  c0.Tags = {{"role", "node"}, {"rc_tm", "100ms"}};
  c0.ReconnectTimeoutOverride = reconnect_override;
  c1.NodeName = "client1";
  c1.Tags = {{"role", "node"}, {"rc_tm", "100ms"}};
  c1.ReconnectTimeoutOverride = reconnect_override;
  auto srv = Serf::Create(pool, cs), cl0 = Serf::Create(pool, c0), cl1 = Serf::Create(pool, c1);
  cl0->Join({"server/x"}, true);
  cl1->Join({"server/x"}, true);
  CHECK(eventually(pool, 140, [&] { return srv->Members().size() == 3 && cl0->Members().size() == 3; }));
  cl1->Shutdown();
  // 1 s of simulated time (20 ticks of 50 ms) after the failure is detected is the test's allowance
  CHECK(eventually(pool, 400, [&] { return count_status(*srv, StatusFailed) == 1 || srv->Members().size() == 2; }));
  CHECK(eventually(pool, 20, [&] { return srv->Members().size() == 2 && cl0->Members().size() == 2; }));
  std::puts("PASS TestClient_ShortReconnectTimeout");
}

// GetCoordinate / GetCachedCoordinate (agent/router/router.go:62-67 sorts servers by the distance
// between these): two datacenters 1 s apart end up about 1 s apart in coordinate space.
static void test_coordinates() {
  gsim_config c = Pool::DefaultWANConfig();
  c.capacity = 256;
  c.n_initial = 254;  // two tiles of anonymous members = two datacenters, plus two named agents
  c.seed = 9;
  c.flags = GSIM_FLAG_COORDINATES;
  c.mailbox_depth = 8;
  Pool pool(c);
  const uint8_t lat[4] = {1, 2, 2, 1};  // one tick of extra latency each way between the datacenters
  CHECK(gsim_latency_set(pool.handle(), 2, lat) == 0);
  Config ca, cb;
  ca.NodeName = "a.dc2";
  cb.NodeName = "b.dc2";
  auto a = Serf::Create(pool, ca), b = Serf::Create(pool, cb);  // ids 254, 255: second tile = datacenter 1
  uint32_t seed = 0;
  int n_ok = 0;
  CHECK(gsim_join(pool.handle(), a->id(), &seed, 1, 1, &n_ok) == 0 && n_ok == 1);
  CHECK(gsim_join(pool.handle(), b->id(), &seed, 1, 1, &n_ok) == 0 && n_ok == 1);
  pool.Step(3000);
  Coordinate here = a->GetCoordinate(), near, far;
  CHECK(a->GetCachedCoordinate("b.dc2", &near) && !a->GetCachedCoordinate("nosuch", &near));
  double w[11];
  CHECK(gsim_coordinate_get(pool.handle(), 3, w) == 0);  // member 3 lives in the other datacenter
  for (int k = 0; k < 8; ++k) far.Vec[k] = w[k];
  far.Error = w[8];
  far.Adjustment = w[9];
  far.Height = w[10];
  CHECK(here.DistanceTo(near) < 0.2);                                  // same datacenter: sub-tick
  CHECK(here.DistanceTo(far) > 0.6 && here.DistanceTo(far) < 1.4);     // 2 x 0.5 s of extra latency
  std::puts("PASS GetCoordinate / GetCachedCoordinate");
}

static void test_user_event() {
  Pool pool(test_cfg());
  std::deque<Event> chs, chc;
  Config cs, cc;
  cs.NodeName = "server";
  cs.EventCh = &chs;
  cc.NodeName = "client";
  cc.EventCh = &chc;
  auto srv = Serf::Create(pool, cs), cli = Serf::Create(pool, cc);
  cli->Join({"server/x"}, true);
  CHECK(eventually(pool, 140, [&] { return srv->NumNodes() == 2 && cli->NumNodes() == 2; }));
  srv->UserEvent("consul:event:foo", "bar", false);
  pool.Step(60);
  pool.PumpEvents();
  int got_s = 0, got_c = 0;
  for (auto& e : chs) got_s += e.Type == EventUser && e.Name == "consul:event:foo" && e.Payload == "bar";
  for (auto& e : chc) got_c += e.Type == EventUser && e.Name == "consul:event:foo" && e.Payload == "bar";
  CHECK(got_s == 1 && got_c == 1);  // exactly once on both
  bool threw = false;
  try {
    srv->UserEvent(std::string(400, 'n'), std::string(400, 'p'), false);  // > UserEventSizeLimit
  } catch (const Error& e) {
    threw = e.code == GSIM_ERR_TOO_LARGE;
  }
  CHECK(threw);
  std::puts("PASS TestClientServer_UserEvent");
}

static void test_leave() {
  gsim_config c = Pool::DefaultLANConfig();
  c.capacity = 64;
  c.n_initial = 40;
  Pool pool(c);
  Config ca;
  ca.NodeName = "a1";
  auto a = Serf::Create(pool, ca);
  // the 40 pre-converged members are anonymous; join through member 0 by id name
  int rc = 0;
  uint32_t seed = 0;
  int n_ok = 0;
  rc = gsim_join(pool.handle(), a->id(), &seed, 1, 1, &n_ok);
  CHECK(rc == 0 && n_ok == 1);
  CHECK(eventually(pool, 200, [&] { return a->NumNodes() == 41; }));
  a->Leave();
  CHECK(a->LocalMember().Status == StatusLeft);
  pool.Step(300);
  CHECK(a->Stats().at("left") == "1");
  std::puts("PASS TestAgent_Leave");
}

// `--extended` runs only the tests added after round 1's GPU verification (reaper timers, WAN);
// without arguments the original set runs, unchanged.
int main(int argc, char** argv) {
  const bool extended = argc > 1 && std::string(argv[1]) == "--extended";
  try {
    if (extended) {
      test_lan_reap_timers();
      test_join_wan();
      test_wan_reap();
      test_set_tags();
      test_merge_delegate();
      test_short_reconnect_timeout();
      test_coordinates();
      std::puts("ALL PASS");
      return 0;
    }
    test_join_lan();
    test_lan_reap_and_force_leave();
    test_user_event();
    test_leave();
  } catch (const Error& e) {
    std::fprintf(stderr, "gsim error %d: %s\n", e.code, e.what());
    return 2;
  }
  std::puts("ALL PASS");
  return 0;
}
