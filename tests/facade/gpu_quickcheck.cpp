// gpu_quickcheck.cpp — TEST INFRASTRUCTURE.  A torch-free, seconds-long parity check of libgsim.so
// (the CUDA path through the C ABI) against the oracle (oracle/liboracle.so) on the features that
// were added after round 1's GPU verification: WAN latency ring, slow links with loss, periodic
// push-pull, the reaper, SetTags, the CSR peer graph (including attaching it after CUDA graphs
// were captured), plus one plain LAN scenario as a regression check of the shared tick kernel.
// Every scenario compares the 256-bit state digest and the counters after every chunk of ticks.
//
// build: g++ -O1 -std=c++17 -Iinclude tests/facade/gpu_quickcheck.cpp -Lconsul_b200 -lgsim
//            -Loracle -loracle -Wl,-rpath,$PWD/consul_b200 -Wl,-rpath,$PWD/oracle
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gsim.h"

extern "C" {
void* oracle_create(const gsim_config* cfg, int threads);
void oracle_destroy(void* h);
int oracle_member_add(void* h, const gsim_member_desc* desc, uint32_t* id_out);
int oracle_join(void* h, uint32_t id, const uint32_t* seeds, size_t n_seeds, int ignore_old, int* n_ok);
int oracle_crash_many(void* h, const uint32_t* ids, size_t n);
int oracle_leave(void* h, uint32_t id);
int oracle_user_event(void* h, uint32_t id, const void* name, size_t nl, const void* payload, size_t pl, int coalesce,
                      uint32_t* slot_out);
int oracle_member_update(void* h, uint32_t id, uint32_t alive_msg_size, uint32_t* slot_out);
int oracle_latency_set(void* h, uint32_t n_dcs, const uint8_t* lat);
int oracle_graph_set(void* h, uint32_t n_rows, const uint32_t* row_ptr, const uint32_t* col_idx);
int oracle_step(void* h, uint32_t ticks);
int oracle_stats_get(void* h, gsim_stats* out);
int oracle_state_hash(void* h, uint64_t out[4]);
}

static int failures = 0;

struct Pair {
  gsim_pool* g = nullptr;
  void* o = nullptr;
  std::string name;
  Pair(const char* nm, const gsim_config& c) : name(nm) {
    int rc = gsim_pool_create(&c, &g);
    if (rc) {
      std::fprintf(stderr, "FAIL %s: gsim_pool_create -> %d (%s)\n", nm, rc, gsim_strerror(rc));
      std::exit(3);
    }
    o = oracle_create(&c, 0);
  }
  ~Pair() {
    gsim_pool_destroy(g);
    oracle_destroy(o);
  }
  void both(int a, int b, const char* what) {
    if (a != b) {
      std::fprintf(stderr, "FAIL %s: %s returned %d vs oracle %d\n", name.c_str(), what, a, b);
      ++failures;
    }
  }
  void add() {
    uint32_t x = 0, y = 0;
    gsim_member_desc d = {0, 0};
    both(gsim_member_add(g, &d, &x), oracle_member_add(o, &d, &y), "member_add");
    both((int)x, (int)y, "member id");
  }
  void join(uint32_t id, uint32_t seed) {
    int a = 0, b = 0;
    both(gsim_join(g, id, &seed, 1, 1, &a), oracle_join(o, id, &seed, 1, 1, &b), "join");
    both(a, b, "join n_ok");
  }
  void crash(std::vector<uint32_t> ids) {
    both(gsim_crash_many(g, ids.data(), ids.size()), oracle_crash_many(o, ids.data(), ids.size()), "crash_many");
  }
  void leave(uint32_t id) { both(gsim_leave(g, id), oracle_leave(o, id), "leave"); }
  void event(uint32_t id, const char* nm, size_t payload) {
    std::string p(payload, 'x');
    uint32_t a = 0, b = 0;
    both(gsim_user_event(g, id, nm, std::strlen(nm), p.data(), p.size(), 0, &a),
         oracle_user_event(o, id, nm, std::strlen(nm), p.data(), p.size(), 0, &b), "user_event");
    both((int)a, (int)b, "event slot");
  }
  void update(uint32_t id) {
    uint32_t a = 0, b = 0;
    both(gsim_member_update(g, id, 120, &a), oracle_member_update(o, id, 120, &b), "member_update");
  }
  void latency(uint32_t n_dcs, bool slow, uint32_t worst) {
    std::vector<uint8_t> m(n_dcs * n_dcs);
    for (uint32_t a = 0; a < n_dcs; ++a)
      for (uint32_t b = 0; b < n_dcs; ++b)
        m[a * n_dcs + b] = a == b ? 1 : (uint8_t)(1 + (slow ? (3 * a + 5 * b) % worst : (7 * a + 13 * b) % 5));
    both(gsim_latency_set(g, n_dcs, m.data()), oracle_latency_set(o, n_dcs, m.data()), "latency_set");
  }
  void graph(const std::vector<uint32_t>& rp, const std::vector<uint32_t>& ci) {
    const uint32_t rows = rp.empty() ? 0 : (uint32_t)rp.size() - 1;
    both(gsim_graph_set(g, rows, rows ? rp.data() : nullptr, rows ? ci.data() : nullptr),
         oracle_graph_set(o, rows, rows ? rp.data() : nullptr, rows ? ci.data() : nullptr), "graph_set");
  }
  void step(uint32_t ticks, uint32_t every) {
    for (uint32_t done = 0; done < ticks;) {
      const uint32_t k = ticks - done < every ? ticks - done : every;
      both(gsim_step(g, k), oracle_step(o, k), "step");
      done += k;
      uint64_t hg[4], ho[4];
      gsim_state_hash(g, hg);
      oracle_state_hash(o, ho);
      gsim_stats sg, so;
      gsim_stats_get(g, &sg);
      oracle_stats_get(o, &so);
      bool same = std::memcmp(hg, ho, sizeof(hg)) == 0;
      for (int q = 0; q < GSIM_STAT_COUNT; ++q)
        if (q != GSIM_STAT_ACTIVE_ROWS && sg.counters[q] != so.counters[q]) same = false;
      if (!same) {
        std::fprintf(stderr, "FAIL %s: diverged by tick %u: digest %016llx vs %016llx\n", name.c_str(), gsim_now(g),
                     (unsigned long long)hg[0], (unsigned long long)ho[0]);
        for (int q = 0; q < GSIM_STAT_COUNT; ++q)
          if (sg.counters[q] != so.counters[q])
            std::fprintf(stderr, "   counter %d: %llu vs %llu\n", q, (unsigned long long)sg.counters[q],
                         (unsigned long long)so.counters[q]);
        ++failures;
        return;
      }
    }
  }
  void done() {
    uint64_t hg[4];
    gsim_state_hash(g, hg);
    std::printf("%s %s  tick %u digest %016llx\n", failures ? "----" : "PASS", name.c_str(), gsim_now(g),
                (unsigned long long)hg[0]);
    std::fflush(stdout);
  }
};

static gsim_config lan(uint32_t n, uint64_t seed, uint32_t extra_cap = 4) {
  gsim_config c;
  gsim_config_default_lan(&c);
  c.capacity = n + extra_cap;
  c.n_initial = n;
  c.seed = seed;
  return c;
}
static gsim_config wan(uint32_t n, uint64_t seed) {
  gsim_config c;
  gsim_config_default_wan(&c);
  c.capacity = n + 4;
  c.n_initial = n;
  c.seed = seed;
  c.mailbox_depth = 8;
  return c;
}

static void ring_of_segments(uint32_t n, uint32_t seg, std::vector<uint32_t>& rp, std::vector<uint32_t>& ci) {
  const uint32_t nseg = (n + seg - 1) / seg;
  rp.assign(1, 0);
  ci.clear();
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t s = i / seg, t = (s + 1) % nseg;
    const uint32_t lo = s < t ? s : t, hi = s < t ? t : s;
    for (uint32_t c = lo * seg; c < (lo + 1) * seg && c < n; ++c) ci.push_back(c);
    if (hi != lo)
      for (uint32_t c = hi * seg; c < (hi + 1) * seg && c < n; ++c) ci.push_back(c);
    rp.push_back((uint32_t)ci.size());
  }
}

int main() {
  {  // the shared tick kernel on the plain LAN path (regression)
    Pair p("lan_join_event_crash_20k", lan(20000, 0x5EED0001));
    p.add();
    p.join(20000, 0);
    p.event(5, "deploy", 32);
    p.crash({10, 11, 12});
    p.step(320, 64);
    p.done();
  }
  {  // BASELINE config 5's latency matrix over a deeper mailbox ring
    Pair p("wan_c5_matrix_event_16k", wan(16461, 0x5EED0005));
    p.latency(64, false, 5);
    p.event(0, "deploy", 32);
    p.step(160, 16);
    p.done();
  }
  {  // links slower than ProbeTimeout, 20 % loss, no TCP fallback: late acks, budgets, refutes
    gsim_config c = wan(3000, 21);
    c.packet_loss_ppm = 200000;
    c.disable_tcp_pings = 1;
    Pair p("wan_slow_links_lossy_3k", c);
    p.latency(16, true, 7);
    p.add();
    p.join(3000, 3);
    p.event(5, "e1", 7);
    p.crash({10, 300, 1200});
    p.step(700, 50);
    p.done();
  }
  {  // periodic push-pull completing a stranded broadcast
    gsim_config c = lan(3000, 0x5EED00AA);
    c.flags = GSIM_FLAG_PUSH_PULL;
    c.push_pull_interval_ns = 1000000000ull;
    c.packet_loss_ppm = 500000;
    c.retransmit_mult = 1;
    Pair p("pushpull_stranded_3k", c);
    p.event(17, "deploy", 2);
    p.add();
    p.join(3000, 3);
    p.crash({100, 200});
    p.step(900, 50);
    p.done();
  }
  {  // serf's reaper with TestServer_LANReap timings
    gsim_config c;
    gsim_config_consul_test(&c);
    c.capacity = 8;
    c.n_initial = 0;
    c.seed = 1;
    c.phase_group = 1;
    c.flags = GSIM_FLAG_LOG_GLOBAL_EVENTS;
    c.reconnect_timeout_ns = 250000000ull;
    c.tombstone_timeout_ns = 250000000ull;
    c.reap_interval_ns = 300000000ull;
    Pair p("lan_reap_3", c);
    p.add();
    p.add();
    p.add();
    p.join(1, 0);
    p.join(2, 0);
    p.step(40, 1);
    p.crash({1});
    p.step(80, 1);
    p.leave(2);
    p.step(120, 4);
    p.done();
  }
  {  // SetTags -> next incarnation, EventMemberUpdate
    gsim_config c = lan(500, 41);
    c.packet_loss_ppm = 200000;
    Pair p("set_tags_500", c);
    p.step(7, 7);
    p.update(5);
    p.step(90, 10);
    p.update(5);
    p.step(60, 10);
    p.done();
  }
  {  // CSR peer graph attached after 64-tick CUDA graphs were captured, then detached again
    Pair p("csr_ring_of_segments_1k", lan(1024, 17, 0));
    p.step(200, 100);
    std::vector<uint32_t> rp, ci;
    ring_of_segments(1024, 128, rp, ci);
    p.graph(rp, ci);
    p.event(0, "e", 1);
    p.crash({300, 700});
    p.step(320, 64);
    p.graph({}, {});
    p.step(128, 64);
    p.done();
  }
  if (failures) {
    std::fprintf(stderr, "%d FAILURE(S)\n", failures);
    return 1;
  }
  std::puts("ALL PASS");
  return 0;
}
