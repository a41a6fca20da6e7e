// oracle.cpp — CPU restatement of the gossip hot path.  TEST INFRASTRUCTURE.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
// build, load or call this file.  The product (consul_b200/, libgsim.so) never does.
//
// PARITY UNPINNED: the algorithm restated here lives in two Go modules that are NOT present
// in /root/reference (github.com/hashicorp/memberlist v0.5.2 and github.com/hashicorp/serf
// v0.10.2, /root/reference/go.mod:80,85; hashes go.sum:528-529,545-546), there is no Go
// toolchain in this image, and no test or fixture under /root/reference asserts an
// incarnation number, Lamport time, suspicion timeout or convergence tick.
// What pins this file: (1) the known-answer values of the upstream formulas carried in
// tests/test_oracle_kats.py, (2) the defaults and enums that Consul's own files pin
// (agent/config/runtime.go:1271-1413, api/agent.go:299-303, agent/consul/server_test.go:221-237),
// (3) the eventual-outcome scenarios of the reference tests (server_test.go:509-529,
// 666-733; client_test.go:756-835) and the 100k-node design figure of
// internal/gossip/libserf/serf.go:29-33, replayed in tests/test_oracle_scenarios.py, (4) the
// queue-order / event-window / de-dup / leave / refute predicates of SURVEY.md §8c as known-answer
// tests in tests/test_predicate_kats.py, (5) oracle/m0_memberlist.py — a full-fidelity per-observer
// restatement ("M0") whose outcomes, exact counts and detection / dissemination times this
// projected model ("M1") must reproduce (tests/test_m0_crosscheck.py).
//
// Model (documented in DESIGN.md §3): lock-step ticks of tau = gcd(ProbeInterval,
// ProbeTimeout, GossipInterval).  Every member reads the cluster as it was published at the
// start of the tick and writes only itself; what it sends arrives at the next tick.  Exact
// per-observer views are O(N^2), so each member's alive/suspect/dead state is ONE shared
// record per subject (the earliest suspicion timer with all confirmations), while the
// epidemic spread of individual broadcasts is tracked exactly, per member, for up to 31
// concurrent "rumors" (joins, intents, user events).
//
// Written independently of consul_b200/csrc (different data structures: one struct per
// member, explicit published views, explicit message lists) so that agreement between the
// two is evidence, not tautology.  Shared: include/gsim.h (the interface structs only).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/gsim.h"

namespace {

const uint32_t NONE32 = 0xFFFFFFFFu;
enum { ST_IDLE = 0, ST_WAIT_TIMEOUT = 1, ST_WAIT_DEADLINE = 2 };
enum { PUR_PHASE = 1, PUR_PERM = 2, PUR_GOSSIP = 3, PUR_RELAY = 4, PUR_LOSS = 5, PUR_CRASH = 6, PUR_PUSHPULL = 7, PUR_COORD = 8 };
const uint32_t PUSHPULL_BACKLOG = 4;  // push-pull connections one member serves per tick
enum { LK_PING = 0, LK_ACK, LK_INDREQ, LK_INDPING, LK_INDACK, LK_INDFWD, LK_NACK, LK_GOSSIP };
const uint32_t KRANDOM_MAX_TRIES = 32;  // upstream: 3n (memberlist/util.go kRandomNodes)
const uint32_t PROBE_SKIP_CAP = 1024;   // upstream: len(nodes)
const int MAX_SUS = GSIM_MAX_SUSPICION_SLOTS;

// width in bits of the probe ring's permutation domain: the smallest power of two >= n (at least 4)
uint32_t ring_bits(uint32_t n) {
  uint32_t bits = 0;
  while (bits < 32 && (1ull << bits) < n) ++bits;
  return std::max(bits, 2u);
}

// ---- Philox4x32-10, written from the Random123 specification ----------------------
struct Rand4 {
  uint32_t v[4];
};
Rand4 philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
  uint32_t ctr[4] = {c0, c1, c2, c3};
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
    uint32_t n[4];
    n[0] = (uint32_t)(p1 >> 32) ^ ctr[1] ^ key0;
    n[1] = (uint32_t)p1;
    n[2] = (uint32_t)(p0 >> 32) ^ ctr[3] ^ key1;
    n[3] = (uint32_t)p0;
    memcpy(ctr, n, sizeof(n));
    key0 += 0x9E3779B9u;  // golden ratio
    key1 += 0xBB67AE85u;  // sqrt(3) - 1
  }
  Rand4 r;
  memcpy(r.v, ctr, sizeof(ctr));
  return r;
}

// ---- ticker stagger: one draw per phase group ([U] state.go triggerFunc: rand % interval) ----
// murmur3's 32-bit finaliser over (seed, group); multiply-high maps it onto [0, interval).
uint32_t murmur_fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
struct Stagger {
  uint32_t probe, gossip;
};
// Phases are dealt round-robin over consecutive phase groups, rotated by a seed-derived offset:
// group g probes at ticks = g + rot (mod P) and gossips at ticks = g / P + rot' (mod GI).
Stagger stagger_of(uint64_t seed, uint32_t member, uint32_t phase_group, uint32_t P, uint32_t GI) {
  uint64_t group = member / (phase_group ? phase_group : 128);
  uint32_t rot = murmur_fmix32((uint32_t)seed * 0x9E3779B1u + (uint32_t)(seed >> 32));
  Stagger s;
  s.probe = (uint32_t)((group + rot % P) % P);
  s.gossip = (uint32_t)((group / P + (rot >> 16) % GI) % GI);
  return s;
}

// ---- upstream scalar formulas ---------------------------------------------------------
// [U] memberlist/util.go retransmitLimit
uint32_t retransmit_limit(uint32_t mult, uint32_t n) {
  return mult * (uint32_t)ceil(log10((double)(n + 1.0)));
}
// [U] memberlist/util.go suspicionTimeout
uint64_t suspicion_timeout_ns(uint32_t mult, uint32_t n, uint64_t interval_ns) {
  double scale = log10(n < 1 ? 1.0 : (double)n);
  if (scale < 1.0) scale = 1.0;
  int64_t milli = (int64_t)(scale * 1000.0);
  return (uint64_t)((int64_t)mult * milli * (int64_t)interval_ns / 1000);
}
// [U] memberlist/suspicion.go remainingSuspicionTime (total timeout; caller subtracts elapsed)
uint64_t suspicion_total_ns(uint32_t confirmations, uint32_t k, uint64_t min_ns, uint64_t max_ns) {
  if (k < 1) return min_ns;
  double frac = log(confirmations + 1.0) / log(k + 1.0);
  double maxs = max_ns / 1e9, mins = min_ns / 1e9;
  double raw = maxs - frac * (maxs - mins);
  int64_t t = (int64_t)floor(1000.0 * raw) * 1000000;
  return t < (int64_t)min_ns ? min_ns : (uint64_t)t;
}
// [U] memberlist/util.go pushPullScale
uint64_t push_pull_scale_ns(uint64_t interval_ns, uint32_t n) {
  if (n <= 32) return interval_ns;
  double m = ceil(log2((double)n) - log2(32.0)) + 1.0;
  return (uint64_t)m * interval_ns;
}
// [U] serf/lamport.go Witness
uint32_t lamport_witness(uint32_t clock, uint32_t seen) { return seen >= clock ? seen + 1 : clock; }
// [U] memberlist/state.go refute
uint32_t refute_incarnation(uint32_t cur, uint32_t accused) {
  uint32_t inc = cur + 1;                     // nextIncarnation
  if (accused >= inc) inc += accused - inc + 1;  // skipIncarnation
  return inc;
}

// ---- network coordinates ([U] serf/coordinate: Vivaldi + height + adjustment + gravity) ------------
// Restated from the published algorithm (client.go Update = updateVivaldi, updateAdjustment,
// updateGravity; coordinate.go DistanceTo, ApplyForce, unitVectorAt), minus the per-peer latency
// filter.  Plain doubles in the textual order of the formulas; built with -ffp-contract=off.
struct Coordinate {
  double v[8];
  double err, adj, h;
};
const int ADJ_WINDOW = 20;
const double V_ERR_MAX = 1.5, V_CE = 0.25, V_CC = 0.25, V_HEIGHT_MIN = 10.0e-6, V_RHO = 150.0, V_ZERO = 1.0e-6;

Coordinate origin_coordinate() {
  Coordinate c;
  for (double& x : c.v) x = 0.0;
  c.err = V_ERR_MAX;
  c.adj = 0.0;
  c.h = V_HEIGHT_MIN;
  return c;
}
double norm8(const double* v) {
  double acc = 0.0;
  for (int k = 0; k < 8; ++k) acc += v[k] * v[k];
  return sqrt(acc);
}
double raw_distance(const Coordinate& a, const Coordinate& b) {
  double d[8];
  for (int k = 0; k < 8; ++k) d[k] = a.v[k] - b.v[k];
  return norm8(d) + a.h + b.h;
}
double distance_seconds(const Coordinate& a, const Coordinate& b) {  // DistanceTo(...).Seconds()
  double dist = raw_distance(a, b);
  double with_adjustments = dist + a.adj + b.adj;
  if (with_adjustments > 0.0) dist = with_adjustments;
  int64_t nanos = (int64_t)(dist * 1.0e9);  // time.Duration truncates
  return (double)(nanos / 1000000000) + (double)(nanos % 1000000000) / 1.0e9;
}
void apply_force(Coordinate& c, double force, const Coordinate& other, uint64_t seed, uint32_t member, uint32_t tick,
                 uint32_t which) {
  double u[8];
  for (int k = 0; k < 8; ++k) u[k] = c.v[k] - other.v[k];
  double mag = norm8(u);
  if (mag > V_ZERO) {
    double inv = 1.0 / mag;
    for (int k = 0; k < 8; ++k) u[k] = u[k] * inv;
  } else {  // coincident points: a pseudo-random direction
    Rand4 a = philox4x32_10(seed, member, tick, PUR_COORD, which * 2), b = philox4x32_10(seed, member, tick, PUR_COORD, which * 2 + 1);
    for (int k = 0; k < 4; ++k) {
      u[k] = (double)a.v[k] / 4294967296.0 - 0.5;
      u[4 + k] = (double)b.v[k] / 4294967296.0 - 0.5;
    }
    double m2 = norm8(u);
    if (m2 > V_ZERO) {
      double inv = 1.0 / m2;
      for (int k = 0; k < 8; ++k) u[k] = u[k] * inv;
    } else {
      for (int k = 0; k < 8; ++k) u[k] = 0.0;
      u[0] = 1.0;
    }
    mag = 0.0;
  }
  for (int k = 0; k < 8; ++k) c.v[k] = c.v[k] + u[k] * force;
  if (mag > V_ZERO) {
    c.h = (c.h + other.h) * force / mag + c.h;
    if (!(c.h > V_HEIGHT_MIN)) c.h = V_HEIGHT_MIN;
  }
}
bool coordinate_is_valid(const Coordinate& c) {
  bool ok = std::isfinite(c.err) && std::isfinite(c.adj) && std::isfinite(c.h);
  for (double x : c.v) ok = ok && std::isfinite(x);
  return ok;
}

// ---- data ---------------------------------------------------------------------------------
struct View {  // what everybody else can read about a member during a tick
  uint8_t truth = 0, rank = 0, pending = 0;
  uint32_t inc = 0;
};

struct Member {
  View v;  // the member's own, current copy (published to `pub` at the end of the tick)
  bool leaving = false, watched = false, isolated = false;
  uint8_t awareness = 0, stage = ST_IDLE, nack_misses = 0;
  uint32_t gossip_phase = 0;
  uint32_t due = 0;
  uint32_t cursor = 0, pass = 0, probe_target = 0, probe_inc = 0;
  uint32_t sus_start = 0;
  uint32_t sus_from[MAX_SUS];
  uint32_t n_sus_from = 0;
  uint32_t change_tick = 0;
  uint64_t own_reconnect_timeout_ns = 0;  // ReconnectTimeoutOverride result for this member (0: pool default)
  Coordinate coord = origin_coordinate();  // serf's coordinate client of this member
  double adj_samples[ADJ_WINDOW];
  uint32_t adj_index = 0;
  uint32_t ltime_member = 1, ltime_event = 1, event_min = 0;
  uint32_t heard = 0, queued = 0;
  uint8_t tx[GSIM_MAX_RUMORS];
  Member() {
    for (double& x : adj_samples) x = 0.0;
    memset(tx, 0, sizeof(tx));
    for (int i = 0; i < MAX_SUS; ++i) sus_from[i] = NONE32;
  }
};

struct Accusation {
  uint32_t subject, inc, from;
};

// One side of a periodic push-pull exchange ([U] memberlist/state.go pushPullNode, net.go
// sendAndReceiveState): the tracked-broadcast mask travels through the inbox like any packet;
// this record carries the sender's Lamport clocks and, for a request, who to answer.
struct PushPull {
  uint32_t to, from, clock_member, clock_event;
  bool request;
};

struct Rumor {
  uint32_t kind = 0, subject = 0, inc = 0, ltime = 0, origin = 0, size = 0, qclass = 0, start = 0;
  uint32_t heard_count = 0, converged_tick = NONE32;
  std::string name, payload;
};

struct Scheduled {
  uint32_t tick, id;
};

struct Tally {
  uint64_t c[GSIM_STAT_COUNT];
  uint32_t heard[GSIM_MAX_RUMORS];
  std::vector<Accusation> accusations;
  std::vector<gsim_event> events;
  std::vector<std::pair<uint32_t, View>> published;
  std::vector<uint32_t> moved;  // members whose coordinate changed in this tick
  std::vector<PushPull> pushpulls;
  int32_t crashed_dead = 0;
  Tally() { clear(); }
  void clear() {
    memset(c, 0, sizeof(c));
    memset(heard, 0, sizeof(heard));
    accusations.clear();
    events.clear();
    published.clear();
    moved.clear();
    pushpulls.clear();
    crashed_dead = 0;
  }
};

struct Oracle {
  gsim_config cfg;
  uint64_t tick_ns = 0;
  uint32_t now = 0;
  uint32_t P = 0, T = 0, GI = 0, gtd = 0, udp_avail = 0, loss_thr = 0;
  uint32_t limit = 0, sus_k = 0, sus_ticks[MAX_SUS], perm_bits = 2;
  uint32_t up_count = 0;
  uint32_t established = 0;  // members folded into the base set
  std::vector<Member> m;
  std::vector<std::pair<uint32_t, uint32_t>> name_lens;  // (member, node-name bytes) where not the canonical "node-<id>"
  std::vector<View> pub;                 // published views (state at the start of the tick)
  std::vector<uint32_t> pub_change_tick; // published change ticks
  std::vector<Coordinate> pub_coord;     // published coordinates (GSIM_FLAG_COORDINATES), as of the tick start
  static const uint32_t RING = 8;        // arrival slots kept per member (one-way latency <= 7 ticks)
  std::vector<uint32_t> inbox[RING];     // rumor bits by arrival tick mod RING (bit 31: accused)
  // WAN pools (BASELINE config 5): one-way latency in ticks between synthetic datacenters;
  // member i lives in datacenter (i / 128) % n_dcs.  n_dcs == 0: one tick everywhere.
  uint32_t n_dcs = 0;
  uint8_t latency[64][64];
  std::vector<uint32_t> wake;            // 0 = look every tick, else the only tick worth a look
  std::vector<Tally> tallies;            // per-thread accumulators, reused every tick
  std::vector<Accusation> arriving;      // accusations that arrive at the tick about to run
  std::vector<PushPull> pp_arriving;     // push-pull requests / answers arriving at that tick
  uint32_t pp_every = 0, pp_rot = 0;     // push-pull ticker period in ticks (0 = off) and rotation
  // restricted topology: member i only knows adjacency[i] (empty vector of vectors = everyone
  // knows everyone, the converged memberlist)
  std::vector<std::vector<uint32_t>> adjacency;
  Rumor rumor[GSIM_MAX_RUMORS];
  uint32_t active = 0;
  std::vector<Scheduled> shutdowns;
  uint64_t stats[GSIM_STAT_COUNT];
  uint64_t node_ticks = 0;
  uint32_t crashed_alive = 0, crashed_dead_tick = NONE32;
  std::vector<gsim_event> events;
  uint32_t events_dropped = 0;
  uint32_t evcap = 65536;
  int threads = 1;
  std::string err;
};

uint64_t gcd_u64(uint64_t a, uint64_t b) { return b ? gcd_u64(b, a % b) : a; }
uint32_t to_ticks_ceil(uint64_t ns, uint64_t tick) { return (uint32_t)((ns + tick - 1) / tick); }

void retune(Oracle& o) {
  uint32_t n = (uint32_t)o.m.size();
  o.limit = std::min<uint32_t>(255, retransmit_limit(o.cfg.retransmit_mult, n));
  int k = (int)o.cfg.suspicion_mult - 2;  // [U] state.go suspectNode
  if ((int)n - 2 < k) k = 0;
  k = std::max(0, std::min(k, MAX_SUS - 1));
  o.sus_k = (uint32_t)k;
  uint64_t mn = suspicion_timeout_ns(o.cfg.suspicion_mult, n, o.cfg.probe_interval_ns);
  uint64_t mx = (uint64_t)o.cfg.suspicion_max_timeout_mult * mn;
  for (int c = 0; c < MAX_SUS; ++c)
    o.sus_ticks[c] = to_ticks_ceil(suspicion_total_ns(std::min<uint32_t>(c, o.sus_k), o.sus_k, mn, mx), o.tick_ns);
  o.perm_bits = ring_bits(n);
  // [U] state.go schedule: push-pull every pushPullScale(PushPullInterval, n)
  o.pp_every = 0;
  if ((o.cfg.flags & GSIM_FLAG_PUSH_PULL) && o.cfg.push_pull_interval_ns) {
    o.pp_every = std::max(2u, to_ticks_ceil(push_pull_scale_ns(o.cfg.push_pull_interval_ns, n), o.tick_ns));
    uint32_t rot = murmur_fmix32((uint32_t)o.cfg.seed * 0x9E3779B1u + (uint32_t)(o.cfg.seed >> 32));
    o.pp_rot = (rot >> 8) % o.pp_every;
  }
}

// does member i's push-pull ticker fire at tick t?  (phase groups are dealt round-robin)
bool pushpull_due(const Oracle& o, uint32_t i, uint32_t t) {
  if (!o.pp_every) return false;
  const uint32_t group = i / (o.cfg.phase_group ? o.cfg.phase_group : 128);
  return (t + o.pp_rot) % o.pp_every == group % o.pp_every;
}

// The four Feistel round keys of one (member, pass) probe ring.
Rand4 ring_keys(uint64_t seed, uint32_t member, uint32_t pass) {
  const uint32_t lo = (uint32_t)seed, hi = (uint32_t)(seed >> 32);
  Rand4 k;
  k.v[0] = murmur_fmix32(member * 0x9E3779B1u + pass * 0x85EBCA77u + lo);
  k.v[1] = murmur_fmix32(k.v[0] ^ hi ^ 0xC2B2AE3Du);
  k.v[2] = k.v[0] * 0x9E3779B1u + k.v[1];
  k.v[3] = (k.v[1] * 0x85EBCA77u) ^ k.v[0];
  return k;
}

// Feistel permutation of [0,n) with cycle walking — the probe ring of one (member, pass).  The domain is
// the smallest power of two >= n; its two halves differ by one bit when that width is odd (the wider one
// on top), and every round XORs one half with the round function of the other: top, bottom, top, bottom.
uint32_t ring_entry(const Oracle& o, uint32_t position, uint32_t n, const Rand4& keys, uint32_t width = 0) {
  const uint32_t bits = width ? width : o.perm_bits;
  const uint32_t bottom_bits = bits / 2, top_bits = bits - bottom_bits;
  uint32_t x = position;
  for (;;) {
    uint32_t half[2] = {x >> bottom_bits, x & ((1u << bottom_bits) - 1)};   // {top, bottom}
    const uint32_t mask[2] = {(1u << top_bits) - 1, (1u << bottom_bits) - 1};
    for (int round = 0; round < 4; ++round) {
      const int into = round & 1, from = into ^ 1;
      uint32_t f = (half[from] + keys.v[round]) * 0x9E3779B1u;
      f ^= f >> 15;
      f *= 0x85EBCA77u;
      f ^= f >> 13;
      half[into] ^= f & mask[into];
    }
    x = (half[0] << bottom_bits) | half[1];
    if (x < n) return x;
  }
}

bool packet_lost(const Oracle& o, Tally& ta, uint32_t src, uint32_t dst, uint32_t t, uint32_t kind, uint32_t idx) {
  if (!o.loss_thr) return false;
  Rand4 r = philox4x32_10(o.cfg.seed, src, dst, t, PUR_LOSS | (kind << 8) | (idx << 16));
  if (r.v[0] < o.loss_thr) {
    ta.c[GSIM_STAT_PACKETS_LOST]++;
    return true;
  }
  return false;
}

bool packet_lost_silently(const Oracle& o, uint32_t src, uint32_t dst, uint32_t t, uint32_t kind, uint32_t idx) {
  if (!o.loss_thr) return false;
  return philox4x32_10(o.cfg.seed, src, dst, t, PUR_LOSS | (kind << 8) | (idx << 16)).v[0] < o.loss_thr;
}

// one-way latency of a packet in ticks (>= 1)
uint32_t one_way(const Oracle& o, uint32_t src, uint32_t dst) {
  if (!o.n_dcs) return 1;
  return o.latency[(src / 128) % o.n_dcs][(dst / 128) % o.n_dcs];
}
// Round trip beyond the two ticks the lock-step model does not charge to a probe: probes are
// evaluated within their tick, so only latency above one tick per direction delays an ack.
uint32_t round_trip(const Oracle& o, uint32_t a, uint32_t b) { return one_way(o, a, b) + one_way(o, b, a) - 2; }

int alive_slot_of(const Oracle& o, uint32_t subject) {
  for (int r = 0; r < GSIM_MAX_RUMORS; ++r)
    if (((o.active >> r) & 1) && o.rumor[r].kind == GSIM_RUMOR_ALIVE && o.rumor[r].subject == subject) return r;
  return -1;
}

// does member i know that c exists?
bool knows(const Oracle& o, uint32_t i, const Member& me, uint32_t c) {
  if (c == i) return true;
  if (!o.pub[c].pending) return !me.isolated;  // established members: known once joined
  int slot = alive_slot_of(o, c);
  return slot >= 0 && ((me.heard >> slot) & 1);
}

// up to 8 member ids without touching the heap (kRandomNodes never returns more)
struct Picks {
  uint32_t v[8];
  uint32_t n = 0;
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  uint32_t operator[](size_t q) const { return v[q]; }
  bool has(uint32_t c) const {
    for (uint32_t q = 0; q < n; ++q)
      if (v[q] == c) return true;
    return false;
  }
  void push_back(uint32_t c) { v[n++] = c; }
};

// member i's memberlist as an indexable sequence
uint32_t list_length(const Oracle& o, uint32_t i) {
  if (o.adjacency.empty()) return (uint32_t)o.m.size();
  return i < o.adjacency.size() ? (uint32_t)o.adjacency[i].size() : 0;
}
uint32_t list_entry(const Oracle& o, uint32_t i, uint32_t position) {
  return o.adjacency.empty() ? position : o.adjacency[i][position];
}

// [U] memberlist/util.go kRandomNodes
Picks k_random(const Oracle& o, uint32_t i, const Member& me, uint32_t t, uint32_t purpose,
                               uint32_t k, bool relays, uint32_t also_exclude) {
  Picks out;
  const uint32_t n = list_length(o, i);
  if (!n) return out;
  uint64_t tries = std::min<uint64_t>(3ull * n, KRANDOM_MAX_TRIES);
  Rand4 block = {{0, 0, 0, 0}};
  for (uint32_t draw = 0; draw < tries && out.size() < k; ++draw) {
    if (draw % 4 == 0) block = philox4x32_10(o.cfg.seed, i, t, purpose, draw / 4);
    uint32_t c = list_entry(o, i, block.v[draw % 4] % n);  // randomOffset into the member list
    if (c == i || c == also_exclude) continue;
    const View& vc = o.pub[c];
    if (vc.truth == GSIM_TRUTH_NONE) continue;
    if (relays) {
      if (vc.rank != GSIM_RANK_ALIVE) continue;
    } else {
      // [U] state.go gossip(): alive/suspect, or dead for at most GossipToTheDeadTime
      if (vc.rank == GSIM_RANK_LEFT) continue;
      if (vc.rank == GSIM_RANK_DEAD && t - o.pub_change_tick[c] > o.gtd) continue;
    }
    if (!knows(o, i, me, c)) continue;
    if (out.has(c)) continue;
    out.push_back(c);
  }
  return out;
}

// [U] memberlist/queue.go GetBroadcasts + [U] serf/delegate.go GetBroadcasts
uint32_t pick_packet(const Oracle& o, const Member& me) {
  struct Item {
    uint32_t cls, tx, size, slot;
  };
  uint32_t total = 0;
  for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r)
    if ((me.queued >> r) & 1) total += o.rumor[r].size + (o.rumor[r].qclass ? 3 : 2);
  if (total <= o.udp_avail) return me.queued;  // everything fits in one packet
  std::vector<Item> items;
  for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r)
    if ((me.queued >> r) & 1) items.push_back({o.rumor[r].qclass, me.tx[r], o.rumor[r].size, r});
  std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
    if (a.cls != b.cls) return a.cls < b.cls;      // memberlist, then intents, then events
    if (a.tx != b.tx) return a.tx < b.tx;          // fewest transmits first
    if (a.size != b.size) return a.size > b.size;  // longest first
    return a.slot > b.slot;                        // newest first
  });
  uint32_t used = 0, mask = 0;
  for (const Item& it : items) {
    uint32_t overhead = it.cls ? 3 : 2;
    if (used + overhead >= o.udp_avail) continue;
    if (it.size > o.udp_avail - used - overhead) continue;
    mask |= 1u << it.slot;
    used += overhead + it.size;
  }
  return mask;
}

void log_event(const Oracle& o, Tally& ta, uint32_t t, uint32_t type, uint32_t subject, uint32_t observer, uint32_t ltime) {
  (void)o;
  gsim_event e = {t, type, subject, observer, ltime, 0};
  ta.events.push_back(e);
}

// When does this member next need a look if no mail arrives?  0 = every tick (a running
// suspicion timer, a refutation, a non-empty broadcast queue), NONE32 = never.
uint32_t wake_of(const Member& me) {
  if (me.v.truth == GSIM_TRUTH_NONE) return NONE32;
  const bool up = me.v.truth == GSIM_TRUTH_UP;
  if (me.v.rank == GSIM_RANK_SUSPECT) return 0;
  if (up && (me.v.rank != GSIM_RANK_ALIVE || me.queued)) return 0;
  return up ? me.due : NONE32;
}

// One member, one tick.
void member_tick(Oracle& o, uint32_t i, uint32_t t, Tally& ta) {
  Member& me = o.m[i];
  if (me.v.truth == GSIM_TRUTH_NONE) return;
  const bool up = me.v.truth == GSIM_TRUTH_UP;
  const View before = me.v;
  const bool my_gossip_tick = up && (t % o.GI) == me.gossip_phase;
  const uint32_t word = o.inbox[t % Oracle::RING][i];
  o.inbox[t % Oracle::RING][i] = 0;
  const uint32_t inbox = word & 0x7FFFFFFFu;
  // accusations addressed to me (bit 31 was set when they were handed over; the arriving
  // list is sorted by subject)
  const bool accused = (word >> 31) != 0;
  auto lo = o.arriving.end();
  if (accused)
    lo = std::lower_bound(o.arriving.begin(), o.arriving.end(), i,
                          [](const Accusation& a, uint32_t s) { return a.subject < s; });

  const bool my_pushpull_tick = up && pushpull_due(o, i, t);
  if (!inbox && !accused && me.v.rank == GSIM_RANK_ALIVE && !(up && me.due == t) &&
      !(my_gossip_tick && me.queued) && !my_pushpull_tick)
    return;
  // push-pull records addressed to me (sorted by receiver, then sender)
  auto pp_lo = o.pp_arriving.end(), pp_hi = o.pp_arriving.end();
  if (accused && o.pp_every) {
    pp_lo = std::lower_bound(o.pp_arriving.begin(), o.pp_arriving.end(), i,
                             [](const PushPull& a, uint32_t s) { return a.to < s; });
    for (pp_hi = pp_lo; pp_hi != o.pp_arriving.end() && pp_hi->to == i;) ++pp_hi;
  }
  if (up)  // [U] serf/delegate.go MergeRemoteState: clocks first (Witness(remote - 1))
    for (auto a = pp_lo; a != pp_hi; ++a) {
      me.ltime_member = std::max(me.ltime_member, a->clock_member);
      me.ltime_event = std::max(me.ltime_event, a->clock_event);
    }
  // (GSIM_STAT_ACTIVE_ROWS is a scheduling diagnostic of the CUDA implementation; not modelled)

  // -- deliveries ------------------------------------------------------------------------
  if (up) {
    uint32_t fresh = inbox & o.active & ~me.heard;
    for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r) {
      if (!((fresh >> r) & 1)) continue;
      const Rumor& ru = o.rumor[r];
      bool accept = true;
      switch (ru.kind) {
        case GSIM_RUMOR_USER_EVENT:  // [U] serf.handleUserEvent
          me.ltime_event = lamport_witness(me.ltime_event, ru.ltime);
          if (ru.ltime < me.event_min) accept = false;
          else if (me.ltime_event > o.cfg.event_buffer && ru.ltime < me.ltime_event - o.cfg.event_buffer) accept = false;
          if (accept && me.watched) log_event(o, ta, t, GSIM_EVENT_USER, r, i, ru.ltime);
          break;
        case GSIM_RUMOR_JOIN_INTENT:
        case GSIM_RUMOR_LEAVE_INTENT:  // [U] serf.handleNode{Join,Leave}Intent
          me.ltime_member = lamport_witness(me.ltime_member, ru.ltime);
          break;
        case GSIM_RUMOR_ALIVE:  // [U] memberlist.aliveNode (new node) -> serf.handleNodeJoin
          if (me.watched) log_event(o, ta, t, GSIM_EVENT_MEMBER_JOIN, ru.subject, i, 0);
          break;
        case GSIM_RUMOR_UPDATE:  // [U] memberlist.aliveNode (known node, newer incarnation) -> NotifyUpdate
          if (me.watched) log_event(o, ta, t, GSIM_EVENT_MEMBER_UPDATE, ru.subject, i, 0);
          break;
      }
      if (accept) {
        me.heard |= 1u << r;
        me.queued |= 1u << r;
        me.tx[r] = 0;
        ta.heard[r]++;
        ta.c[GSIM_STAT_RUMORS_ACCEPTED]++;
      } else {
        ta.c[GSIM_STAT_RUMORS_DROPPED]++;
      }
    }
  }
  if (accused) {
    // [U] memberlist.suspectNode.  Highest incarnation first, then smallest accuser; at most
    // MAX_SUS per tick (what the mailbox of the CUDA implementation retains).
    int taken = 0;
    for (auto a = lo; a != o.arriving.end() && a->subject == i && taken < MAX_SUS; ++a, ++taken) {
      if (a->inc != me.v.inc) continue;  // stale incarnation
      if (me.v.rank == GSIM_RANK_ALIVE) {
        me.v.rank = GSIM_RANK_SUSPECT;
        me.sus_start = t - 1;
        me.n_sus_from = 1;
        me.sus_from[0] = a->from;
        for (int q = 1; q < MAX_SUS; ++q) me.sus_from[q] = NONE32;
        ta.c[GSIM_STAT_SUSPECTS]++;
      } else if (me.v.rank == GSIM_RANK_SUSPECT) {
        // suspicion.Confirm: a new, distinct accuser; only k confirmations count
        bool seen = false;
        for (uint32_t q = 0; q < me.n_sus_from; ++q) seen |= me.sus_from[q] == a->from;
        if (!seen && me.n_sus_from < o.sus_k + 1) {
          me.sus_from[me.n_sus_from++] = a->from;
          ta.c[GSIM_STAT_CONFIRMATIONS]++;
        }
      }
    }
  }

  // -- push-pull: answer whoever connected ([U] net.go handleConn -> sendLocalState) --------------
  if (up) {
    uint32_t served = 0;
    for (auto a = pp_lo; a != pp_hi && served < PUSHPULL_BACKLOG; ++a) {
      if (!a->request) continue;
      ++served;
      ta.pushpulls.push_back({a->from, i, me.ltime_member, me.ltime_event, false});
      __atomic_fetch_or(&o.inbox[(t + 1) % Oracle::RING][a->from], me.heard & o.active, __ATOMIC_RELAXED);
    }
  }

  // -- my own state as others see it ---------------------------------------------------------
  if (up && !me.leaving && (me.v.rank == GSIM_RANK_SUSPECT || me.v.rank == GSIM_RANK_DEAD)) {
    me.v.inc = refute_incarnation(me.v.inc, me.v.inc);  // [U] memberlist.refute
    me.v.rank = GSIM_RANK_ALIVE;
    me.awareness = (uint8_t)std::min<uint32_t>(me.awareness + 1, o.cfg.awareness_max_multiplier - 1);
    ta.c[GSIM_STAT_REFUTES]++;
  } else if (me.v.rank == GSIM_RANK_SUSPECT) {
    uint32_t confirmations = std::min<uint32_t>(me.n_sus_from - 1, o.sus_k);
    if (t - me.sus_start >= o.sus_ticks[confirmations]) {
      me.v.rank = GSIM_RANK_DEAD;  // [U] memberlist.deadNode
      me.change_tick = t;
      ta.c[GSIM_STAT_DEADS]++;
      if (me.v.truth == GSIM_TRUTH_CRASHED) ta.crashed_dead++;
      if (o.cfg.flags & GSIM_FLAG_LOG_GLOBAL_EVENTS) log_event(o, ta, t, GSIM_EVENT_MEMBER_FAILED, i, NONE32, 0);
    }
  }

  if (up) {
    // -- failure detector ([U] memberlist.probeNode) -------------------------------------------
    if (me.stage == ST_WAIT_TIMEOUT && me.due == t) {
      const uint32_t j = me.probe_target;
      const bool target_up = o.pub[j].truth == GSIM_TRUTH_UP;
      Picks relays = k_random(o, i, me, t, PUR_RELAY, std::min<uint32_t>(8, o.cfg.indirect_checks), true, j);
      bool success = false;
      uint32_t nacks = 0;
      const uint32_t started = t - o.T;
      // ticks left until the probe deadline (started + P * (awareness + 1)); on a WAN pool an
      // answer only counts if its extra latency fits in them
      const uint32_t left = started + o.P * (me.awareness + 1u) - t;
      for (uint32_t q = 0; q < relays.size(); ++q) {
        uint32_t r = relays[q];
        ta.c[GSIM_STAT_INDIRECT_PINGS]++;
        if (o.pub[r].truth != GSIM_TRUTH_UP) continue;                  // relay is down
        if (packet_lost(o, ta, i, r, t, LK_INDREQ, q)) continue;        // request lost
        const uint32_t to_relay_and_back = round_trip(o, i, r);
        const uint32_t relay_to_target = round_trip(o, r, j);
        // the relay gives the target ProbeTimeout to ack, then reports a nack
        bool acked = target_up && !packet_lost(o, ta, r, j, t, LK_INDPING, q) &&
                     !packet_lost(o, ta, j, r, t, LK_INDACK, q) && relay_to_target <= o.T;
        if (acked) {
          if (!packet_lost(o, ta, r, i, t, LK_INDFWD, q) && to_relay_and_back + relay_to_target <= left) success = true;
        } else if (!packet_lost(o, ta, r, i, t, LK_NACK, q) && to_relay_and_back <= left) {
          nacks++;
          ta.c[GSIM_STAT_NACKS]++;
        }
      }
      const uint32_t direct = round_trip(o, i, j);
      if (!o.cfg.disable_tcp_pings && target_up && direct <= left) success = true;  // TCP fallback
      // the direct ack may simply have been slower than ProbeTimeout: it counts until the deadline
      if (o.n_dcs && target_up && direct > o.T && started + direct <= t + left &&
          !packet_lost_silently(o, i, j, started, LK_PING, 0) && !packet_lost_silently(o, j, i, started, LK_ACK, 0))
        success = true;
      if (success) {
        if (me.awareness) me.awareness--;
        me.stage = ST_IDLE;
        me.due = started + o.P;
        ta.c[GSIM_STAT_ACKS]++;
      } else {
        me.nack_misses = (uint8_t)(relays.empty() ? 1 : relays.size() - nacks);
        me.stage = ST_WAIT_DEADLINE;
        me.due = started + o.P * (me.awareness + 1u);
      }
    }
    if (me.stage == ST_WAIT_DEADLINE && me.due == t) {
      me.awareness = (uint8_t)std::min<uint32_t>(me.awareness + me.nack_misses, o.cfg.awareness_max_multiplier - 1);
      me.stage = ST_IDLE;
      ta.accusations.push_back({me.probe_target, me.probe_inc, i});
      ta.c[GSIM_STAT_PROBE_FAILURES]++;
    }
    if (me.stage == ST_IDLE && me.due == t) {
      // [U] memberlist.probe: walk the ring to the next probe-able member
      const uint32_t n = list_length(o, i);
      const uint32_t half_bits = o.adjacency.empty() ? 0 : ring_bits(n);
      Rand4 keys = ring_keys(o.cfg.seed, i, me.pass);
      uint32_t checked = 0, target = NONE32;
      const uint32_t cap = std::min(n, PROBE_SKIP_CAP);
      while (checked < cap) {
        if (me.cursor >= n) {  // wrapped: reshuffle
          me.cursor = 0;
          me.pass++;
          checked++;
          keys = ring_keys(o.cfg.seed, i, me.pass);
          continue;
        }
        uint32_t c = list_entry(o, i, ring_entry(o, me.cursor++, n, keys, half_bits));
        const View& vc = o.pub[c];
        if (c == i || vc.truth == GSIM_TRUTH_NONE || vc.rank == GSIM_RANK_DEAD || vc.rank == GSIM_RANK_LEFT ||
            !knows(o, i, me, c)) {
          checked++;
          continue;
        }
        target = c;
        break;
      }
      if (target == NONE32) {
        me.due = t + o.P;
      } else {
        ta.c[GSIM_STAT_PROBES]++;
        bool acked = o.pub[target].truth == GSIM_TRUTH_UP && !packet_lost(o, ta, i, target, t, LK_PING, 0) &&
                     !packet_lost(o, ta, target, i, t, LK_ACK, 0) && round_trip(o, i, target) <= o.T;
        if (acked) {
          if (me.awareness) me.awareness--;
          me.due = t + o.P;
          ta.c[GSIM_STAT_ACKS]++;
          if (o.cfg.flags & GSIM_FLAG_COORDINATES) {
            // [U] serf/ping_delegate.go NotifyPingComplete -> coordinate.Client.Update(other, rtt)
            const Coordinate& other = o.pub_coord[target];
            double rtt = 0.0005 + (double)round_trip(o, i, target) * ((double)o.tick_ns / 1.0e9);
            Coordinate& c = me.coord;
            double dist = distance_seconds(c, other);                       // updateVivaldi
            if (rtt < V_ZERO) rtt = V_ZERO;
            double wrongness = fabs(dist - rtt) / rtt;
            double total_error = c.err + other.err;
            if (total_error < V_ZERO) total_error = V_ZERO;
            double weight = c.err / total_error;
            c.err = V_CE * weight * wrongness + c.err * (1.0 - V_CE * weight);
            if (c.err > V_ERR_MAX) c.err = V_ERR_MAX;
            double force = V_CC * weight * (rtt - dist);
            apply_force(c, force, other, o.cfg.seed, i, t, 0);
            me.adj_samples[me.adj_index] = rtt - raw_distance(c, other);    // updateAdjustment
            me.adj_index = (me.adj_index + 1) % ADJ_WINDOW;
            double sum = 0.0;
            for (double x : me.adj_samples) sum += x;
            c.adj = sum / (2.0 * (double)ADJ_WINDOW);
            Coordinate org = origin_coordinate();                            // updateGravity
            double q = distance_seconds(org, c) / V_RHO;
            apply_force(c, -1.0 * (q * q), org, o.cfg.seed, i, t, 1);
            if (!coordinate_is_valid(c)) c = origin_coordinate();
            ta.moved.push_back(i);
          }
        } else {
          me.stage = ST_WAIT_TIMEOUT;
          me.probe_target = target;
          me.probe_inc = o.pub[target].inc;
          me.due = t + o.T;
        }
      }
    }

    // -- dissemination ([U] memberlist.gossip) --------------------------------------------------
    if (my_gossip_tick && me.queued) {
      Picks peers = k_random(o, i, me, t, PUR_GOSSIP, std::min<uint32_t>(8, o.cfg.gossip_nodes), false, NONE32);
      for (uint32_t q = 0; q < peers.size() && me.queued; ++q) {
        uint32_t packet = pick_packet(o, me);
        if (!packet) break;
        for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r) {
          if (!((packet >> r) & 1)) continue;
          me.tx[r]++;
          if (me.tx[r] >= o.limit) me.queued &= ~(1u << r);  // retransmit budget spent
          ta.c[GSIM_STAT_RUMORS_SENT]++;
        }
        ta.c[GSIM_STAT_GOSSIP_PACKETS]++;
        if (!packet_lost(o, ta, i, peers[q], t, LK_GOSSIP, q))
          __atomic_fetch_or(&o.inbox[(t + one_way(o, i, peers[q])) % Oracle::RING][peers[q]], packet, __ATOMIC_RELAXED);
      }
    }
  }

  // -- anti-entropy ([U] memberlist.pushPull): one random alive peer, full state both ways --------
  if (my_pushpull_tick && !me.isolated) {
    Picks partner = k_random(o, i, me, t, PUR_PUSHPULL, 1, true, NONE32);
    if (!partner.empty()) {
      ta.pushpulls.push_back({partner[0], i, me.ltime_member, me.ltime_event, true});
      __atomic_fetch_or(&o.inbox[(t + 1) % Oracle::RING][partner[0]], me.heard & o.active, __ATOMIC_RELAXED);
      ta.c[GSIM_STAT_PUSH_PULLS]++;
    }
  }

  if (me.v.inc != before.inc || me.v.rank != before.rank) ta.published.push_back({i, me.v});
}

void run_tick(Oracle& o) {
  const uint32_t t = o.now;
  const uint32_t n = (uint32_t)o.m.size();
  std::sort(o.arriving.begin(), o.arriving.end(), [](const Accusation& a, const Accusation& b) {
    if (a.subject != b.subject) return a.subject < b.subject;
    if (a.inc != b.inc) return a.inc > b.inc;
    return a.from < b.from;
  });
  o.arriving.erase(std::unique(o.arriving.begin(), o.arriving.end(),
                               [](const Accusation& a, const Accusation& b) {
                                 return a.subject == b.subject && a.inc == b.inc && a.from == b.from;
                               }),
                   o.arriving.end());
  for (const Accusation& a : o.arriving) o.inbox[t % Oracle::RING][a.subject] |= 0x80000000u;
  // a member serves the PUSHPULL_BACKLOG smallest requester ids; sort so that they come first
  std::sort(o.pp_arriving.begin(), o.pp_arriving.end(), [](const PushPull& a, const PushPull& b) {
    if (a.to != b.to) return a.to < b.to;
    return a.from < b.from;
  });
  for (const PushPull& a : o.pp_arriving) o.inbox[t % Oracle::RING][a.to] |= 0x80000000u;
  // Members with no mail whose only scheduled action lies at another tick cannot do anything
  // (this is exactly the idle test at the top of member_tick, evaluated from two flat arrays).
  const uint32_t* mail = o.inbox[t % Oracle::RING].data();
  uint32_t* wake = o.wake.data();
  std::vector<Tally>& tallies = o.tallies;
  tallies.resize((size_t)o.threads);
  for (Tally& ta : tallies) ta.clear();
#ifdef _OPENMP
#pragma omp parallel num_threads(o.threads)
  {
    Tally& ta = tallies[(size_t)omp_get_thread_num()];
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      if (mail[i] == 0 && wake[i] != 0 && wake[i] != t && !pushpull_due(o, (uint32_t)i, t)) continue;
      member_tick(o, (uint32_t)i, t, ta);
      wake[i] = wake_of(o.m[(size_t)i]);
    }
  }
#else
  for (uint32_t i = 0; i < n; ++i) {
    if (mail[i] == 0 && wake[i] != 0 && wake[i] != t && !pushpull_due(o, i, t)) continue;
    member_tick(o, i, t, tallies[0]);
    wake[i] = wake_of(o.m[i]);
  }
#endif
  // end of tick: publish, count, hand accusations to the next tick
  o.arriving.clear();
  o.pp_arriving.clear();
  for (Tally& ta : tallies) {
    o.pp_arriving.insert(o.pp_arriving.end(), ta.pushpulls.begin(), ta.pushpulls.end());
    for (int s = 0; s < GSIM_STAT_COUNT; ++s) o.stats[s] += ta.c[s];
    for (auto& pv : ta.published) {
      o.pub[pv.first] = pv.second;
      o.pub_change_tick[pv.first] = o.m[pv.first].change_tick;
    }
    for (uint32_t who : ta.moved) o.pub_coord[who] = o.m[who].coord;
    o.arriving.insert(o.arriving.end(), ta.accusations.begin(), ta.accusations.end());
    for (const gsim_event& e : ta.events) {
      if (o.events.size() < o.evcap) o.events.push_back(e);
      else o.events_dropped++;
    }
    if (ta.crashed_dead) {
      o.crashed_alive -= (uint32_t)ta.crashed_dead;
      if (o.crashed_alive == 0) o.crashed_dead_tick = t;
    }
  }
  for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r) {
    uint32_t c = 0;
    for (Tally& ta : tallies) c += ta.heard[r];
    if (c) {
      o.rumor[r].heard_count += c;
      if (o.rumor[r].heard_count == o.up_count) o.rumor[r].converged_tick = t;
    }
  }
  o.now = t + 1;
  o.node_ticks += n;
}

// ---- between-tick operations ----------------------------------------------------------------
void publish(Oracle& o, uint32_t i) {
  o.pub[i] = o.m[i].v;
  o.pub_change_tick[i] = o.m[i].change_tick;
}

struct Counts {
  uint32_t heard[GSIM_MAX_RUMORS], queued[GSIM_MAX_RUMORS], truth[4], rank[4], crashed_alive, isolated_up;
};
Counts count_all(const Oracle& o) {
  Counts c;
  memset(&c, 0, sizeof(c));
  for (const Member& me : o.m) {
    c.truth[me.v.truth]++;
    if (me.v.truth == GSIM_TRUTH_NONE) continue;
    c.rank[me.v.rank]++;
    if (me.v.truth == GSIM_TRUTH_CRASHED && me.v.rank < GSIM_RANK_DEAD) c.crashed_alive++;
    if (me.v.truth == GSIM_TRUTH_UP && me.isolated) c.isolated_up++;
    if (me.v.truth == GSIM_TRUTH_UP)
      for (int r = 0; r < GSIM_MAX_RUMORS; ++r)
        if ((o.active >> r) & 1) {
          c.heard[r] += (me.heard >> r) & 1;
          c.queued[r] += (me.queued >> r) & 1;
        }
  }
  return c;
}

void after_truth_change(Oracle& o) {
  Counts c = count_all(o);
  o.up_count = c.truth[GSIM_TRUTH_UP];
  o.crashed_alive = c.crashed_alive;
  o.crashed_dead_tick = (c.crashed_alive == 0 && c.truth[GSIM_TRUTH_CRASHED] > 0) ? o.now : NONE32;
  for (int r = 0; r < GSIM_MAX_RUMORS; ++r)
    if ((o.active >> r) & 1) {
      o.rumor[r].heard_count = c.heard[r];
      if (c.heard[r] == o.up_count && o.rumor[r].converged_tick == NONE32) o.rumor[r].converged_tick = o.now;
    }
}

void retire(Oracle& o, uint32_t slot) {
  Rumor& ru = o.rumor[slot];
  if (ru.kind == GSIM_RUMOR_ALIVE) {
    if (o.m[ru.subject].v.pending) o.established++;
    o.m[ru.subject].v.pending = 0;
    publish(o, ru.subject);
  }
  o.active &= ~(1u << slot);
  ru = Rumor();
  for (Member& me : o.m) {
    me.heard &= ~(1u << slot);
    me.queued &= ~(1u << slot);
  }
  for (uint32_t b = 0; b < Oracle::RING; ++b)
    for (uint32_t& w : o.inbox[b]) w &= ~(1u << slot);
}

void auto_retire(Oracle& o) {
  bool any = false;
  for (int r = 0; r < GSIM_MAX_RUMORS; ++r)
    any |= ((o.active >> r) & 1) && o.rumor[r].kind != GSIM_RUMOR_USER_EVENT;
  if (!any) return;
  Counts c = count_all(o);
  for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r)
    if (((o.active >> r) & 1) && o.rumor[r].kind != GSIM_RUMOR_USER_EVENT && c.heard[r] == o.up_count && c.queued[r] == 0) {
      if (o.rumor[r].kind == GSIM_RUMOR_ALIVE && c.isolated_up) continue;  // someone still relies on the bit
      retire(o, r);
    }
}

int free_slot(Oracle& o) {
  for (int attempt = 0; attempt < 2; ++attempt) {
    for (int r = 0; r < GSIM_MAX_RUMORS; ++r)
      if (!((o.active >> r) & 1)) return r;
    if (attempt == 0) auto_retire(o);
  }
  return -1;
}

void start_rumor(Oracle& o, int slot, uint32_t kind, uint32_t subject, uint32_t inc, uint32_t ltime, uint32_t origin,
                 uint32_t size, uint32_t qclass) {
  Rumor& ru = o.rumor[slot];
  ru = Rumor();
  ru.kind = kind;
  ru.subject = subject;
  ru.inc = inc;
  ru.ltime = ltime;
  ru.origin = origin;
  ru.size = size;
  ru.qclass = qclass;
  ru.start = o.now;
  ru.heard_count = 1;
  ru.converged_tick = o.up_count == 1 ? o.now : NONE32;
  o.active |= 1u << slot;
  Member& me = o.m[origin];
  me.heard |= 1u << slot;
  me.queued |= 1u << slot;
  me.tx[slot] = 0;
}

void host_event(Oracle& o, uint32_t type, uint32_t subject, uint32_t observer, uint32_t ltime) {
  gsim_event e = {o.now, type, subject, observer, ltime, 0};
  if (o.events.size() < o.evcap) o.events.push_back(e);
  else o.events_dropped++;
}

// one direction of the join push-pull ([U] memberlist.mergeState, serf MergeRemoteState)
void merge_from(Oracle& o, uint32_t dst, uint32_t src, bool ignore_old) {
  Member& d = o.m[dst];
  const Member& s = o.m[src];
  d.ltime_member = std::max(d.ltime_member, s.ltime_member);
  d.ltime_event = std::max(d.ltime_event, s.ltime_event);
  if (ignore_old && s.ltime_event > d.event_min) d.event_min = s.ltime_event;
  uint32_t fresh = s.heard & ~d.heard & o.active;
  for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r) {
    if (!((fresh >> r) & 1)) continue;
    Rumor& ru = o.rumor[r];
    bool accept = true;
    if (ru.kind == GSIM_RUMOR_USER_EVENT) {
      d.ltime_event = lamport_witness(d.ltime_event, ru.ltime);
      if (ru.ltime < d.event_min) accept = false;
      else if (d.ltime_event > o.cfg.event_buffer && ru.ltime < d.ltime_event - o.cfg.event_buffer) accept = false;
      if (accept && d.watched) host_event(o, GSIM_EVENT_USER, r, dst, ru.ltime);
    } else if (ru.kind == GSIM_RUMOR_JOIN_INTENT || ru.kind == GSIM_RUMOR_LEAVE_INTENT) {
      d.ltime_member = lamport_witness(d.ltime_member, ru.ltime);
    } else if (ru.kind == GSIM_RUMOR_ALIVE) {
      if (d.watched) host_event(o, GSIM_EVENT_MEMBER_JOIN, ru.subject, dst, 0);
    } else if (ru.kind == GSIM_RUMOR_UPDATE) {
      if (d.watched) host_event(o, GSIM_EVENT_MEMBER_UPDATE, ru.subject, dst, 0);
    }
    if (!accept) continue;
    d.heard |= 1u << r;
    d.queued |= 1u << r;
    d.tx[r] = 0;
    ru.heard_count++;
    if (ru.heard_count == o.up_count && ru.converged_tick == NONE32) ru.converged_tick = o.now;
  }
}

void apply_shutdowns(Oracle& o) {
  bool any = false;
  for (size_t x = 0; x < o.shutdowns.size();) {
    if (o.shutdowns[x].tick <= o.now) {
      Member& me = o.m[o.shutdowns[x].id];
      if (me.v.truth == GSIM_TRUTH_UP) {
        me.v.truth = GSIM_TRUTH_GONE;
        publish(o, o.shutdowns[x].id);
        any = true;
      }
      o.shutdowns.erase(o.shutdowns.begin() + x);
    } else {
      ++x;
    }
  }
  if (any) after_truth_change(o);
}

// [U] serf.handleReap: the reaper wakes every ReapInterval and forgets members that have been
// Failed for longer than ReconnectTimeout, or Left for longer than TombstoneTimeout
// (EventMemberReap).  Consul's test timings: agent/consul/server_test.go:675-677.
void reap(Oracle& o) {
  if (!o.cfg.reap_interval_ns) return;
  const uint64_t every = (o.cfg.reap_interval_ns + o.tick_ns - 1) / o.tick_ns;
  if (o.now == 0 || o.now % every != 0) return;
  const uint64_t failed_for = (o.cfg.reconnect_timeout_ns + o.tick_ns - 1) / o.tick_ns;
  const uint64_t left_for = (o.cfg.tombstone_timeout_ns + o.tick_ns - 1) / o.tick_ns;
  bool any = false;
  for (uint32_t i = 0; i < o.m.size(); ++i) {
    Member& me = o.m[i];
    if (me.v.truth == GSIM_TRUTH_NONE || me.v.truth == GSIM_TRUTH_UP) continue;
    uint64_t keep_for;
    if (me.v.rank == GSIM_RANK_DEAD)
      keep_for = me.own_reconnect_timeout_ns ? std::max<uint64_t>(1, (me.own_reconnect_timeout_ns + o.tick_ns - 1) / o.tick_ns)
                                             : failed_for;
    else if (me.v.rank == GSIM_RANK_LEFT) keep_for = left_for;
    else continue;
    if ((uint64_t)(o.now - me.change_tick) <= keep_for) continue;
    if (!me.v.pending) o.established--;
    me.v.truth = GSIM_TRUTH_NONE;
    publish(o, i);
    if (o.cfg.flags & GSIM_FLAG_LOG_GLOBAL_EVENTS) host_event(o, GSIM_EVENT_MEMBER_REAP, i, NONE32, 0);
    any = true;
  }
  if (any) after_truth_change(o);
}

void step(Oracle& o, uint32_t ticks) {
  for (uint32_t k = 0; k < ticks; ++k) {
    apply_shutdowns(o);
    reap(o);
    if (k == 0 || o.wake.size() != o.m.size()) {  // between-tick operations may have changed anyone
      o.wake.resize(o.m.size());
      for (size_t i = 0; i < o.m.size(); ++i) o.wake[i] = wake_of(o.m[i]);
    }
    run_tick(o);
  }
  apply_shutdowns(o);
  auto_retire(o);
}

uint64_t mix(uint64_t h, uint64_t w) {
  h = (h ^ w) * 0xff51afd7ed558ccdull;
  return h ^ (h >> 32);
}

uint32_t pack_key(const View& v) { return (v.inc << 5) | ((uint32_t)v.pending << 4) | ((uint32_t)v.rank << 2) | v.truth; }
uint32_t pack_meta(const Member& me) {
  return me.awareness | ((uint32_t)me.stage << 3) | ((uint32_t)me.nack_misses << 5) | (me.leaving ? 1u << 9 : 0) |
         (me.watched ? 1u << 10 : 0) | (me.isolated ? 1u << 11 : 0) | (me.gossip_phase << 16);
}

}  // namespace

// msgpack sizes as memberlist and serf encode (zero codec.MsgpackHandle: raw strings — fixraw up to 31
// bytes, raw16 after that; there is no str8 without WriteExt).  Arithmetic only: the product has a
// real encoder (csrc/gs_wire.h) and tests/test_wire.py holds the two against each other.
static uint32_t mp_str(size_t n) { return (uint32_t)(n < 32 ? 1 + n : n < 65536 ? 3 + n : 5 + n); }
static uint32_t mp_uint(uint64_t v) { return v < 128 ? 1 : v < 256 ? 2 : v < 65536 ? 3 : v < 4294967296ull ? 5 : 9; }
// A virtual member is called "node-<id>" unless the descriptor gave the length of its real name.
static uint32_t name_len_of(const Oracle& o, uint32_t id) {
  for (const auto& kv : o.name_lens)
    if (kv.first == id) return kv.second;
  uint32_t digits = 1;
  for (uint32_t v = id; v >= 10; v /= 10) ++digits;
  return 5 + digits;
}
// [U] memberlist alive{Incarnation, Node, Addr(4), Port(8301), Meta, Vsn(6)} behind the message-type byte
static uint32_t alive_bytes(const Oracle& o, uint32_t id, uint32_t inc, uint32_t meta_len) {
  return 1 + 1 + (12 + mp_uint(inc)) + (5 + mp_str(name_len_of(o, id))) + (5 + mp_str(4)) + (5 + mp_uint(8301)) +
         (5 + mp_str(meta_len)) + (4 + mp_str(6));
}
// [U] serf messageJoin{LTime, Node} / messageLeave{LTime, Node, Prune} behind the serf-type byte
static uint32_t intent_bytes(const Oracle& o, uint32_t id, bool leave, uint32_t ltime) {
  return 1 + 1 + (6 + mp_uint(ltime)) + (5 + mp_str(name_len_of(o, id))) + (leave ? 6 + 1 : 0);
}

// =============================================================================================
// C interface used by the tests (mirrors include/gsim.h so scenarios read the same)
// =============================================================================================
extern "C" {

uint32_t oracle_retransmit_limit(uint32_t mult, uint32_t n) { return retransmit_limit(mult, n); }
uint64_t oracle_suspicion_timeout_ns(uint32_t mult, uint32_t n, uint64_t interval_ns) {
  return suspicion_timeout_ns(mult, n, interval_ns);
}
int64_t oracle_remaining_suspicion_ns(uint32_t c, uint32_t k, uint64_t elapsed, uint64_t mn, uint64_t mx) {
  return (int64_t)suspicion_total_ns(c, k, mn, mx) - (int64_t)elapsed;
}
uint64_t oracle_push_pull_scale_ns(uint64_t interval, uint32_t n) { return push_pull_scale_ns(interval, n); }
uint32_t oracle_lamport_witness(uint32_t clock, uint32_t v) { return lamport_witness(clock, v); }
uint32_t oracle_refute_incarnation(uint32_t cur, uint32_t accused) { return refute_incarnation(cur, accused); }
uint32_t oracle_ring_entry(uint64_t seed, uint32_t n, uint32_t member, uint32_t pass, uint32_t position) {
  if (n == 0 || position >= n) return NONE32;
  Oracle tmp;
  return ring_entry(tmp, position, n, ring_keys(seed, member, pass), ring_bits(n));
}
void oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  Rand4 r = philox4x32_10(((uint64_t)key[1] << 32) | key[0], ctr[0], ctr[1], ctr[2], ctr[3]);
  memcpy(out, r.v, 16);
}

void* oracle_create(const gsim_config* cfg, int threads) {
  if (!cfg || cfg->struct_size != sizeof(gsim_config)) return nullptr;
  Oracle* o = new Oracle();
  o->cfg = *cfg;
  uint64_t tick = cfg->tick_ns ? cfg->tick_ns
                               : gcd_u64(gcd_u64(cfg->probe_interval_ns, cfg->probe_timeout_ns), cfg->gossip_interval_ns);
  o->tick_ns = tick;
  o->P = (uint32_t)(cfg->probe_interval_ns / tick);
  o->T = (uint32_t)(cfg->probe_timeout_ns / tick);
  o->GI = (uint32_t)(cfg->gossip_interval_ns / tick);
  o->gtd = to_ticks_ceil(cfg->gossip_to_the_dead_ns, tick);
  o->udp_avail = cfg->udp_buffer_size > 2 ? cfg->udp_buffer_size - 2 : 0;
  o->loss_thr = cfg->packet_loss_ppm >= 1000000 ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)cfg->packet_loss_ppm << 32) / 1000000);
  o->evcap = cfg->event_log_capacity ? cfg->event_log_capacity : 65536;
  memset(o->stats, 0, sizeof(o->stats));
#ifdef _OPENMP
  o->threads = threads > 0 ? threads : omp_get_max_threads();
#else
  (void)threads;
  o->threads = 1;
#endif
  o->m.resize(cfg->n_initial);
  o->pub.resize(cfg->n_initial);
  o->pub_change_tick.assign(cfg->n_initial, 0);
  o->pub_coord.assign(cfg->n_initial, origin_coordinate());
  for (uint32_t b = 0; b < Oracle::RING; ++b) o->inbox[b].assign(cfg->n_initial, 0);
  memset(o->latency, 1, sizeof(o->latency));
  o->up_count = cfg->n_initial;
  o->established = cfg->n_initial;
  for (uint32_t i = 0; i < cfg->n_initial; ++i) {
    Member& me = o->m[i];
    me.v.truth = GSIM_TRUTH_UP;
    me.v.rank = GSIM_RANK_ALIVE;
    me.v.inc = 1;
    // ticker stagger ([U] state.go triggerFunc), drawn once per phase group of members
    Stagger ph = stagger_of(cfg->seed, i, cfg->phase_group, o->P, o->GI);
    me.due = ph.probe;
    me.gossip_phase = ph.gossip;
    o->pub[i] = me.v;
  }
  retune(*o);
  return o;
}

void oracle_destroy(void* h) { delete (Oracle*)h; }
int oracle_threads(void* h) { return ((Oracle*)h)->threads; }
int oracle_set_threads(void* h, int threads) {
  Oracle& o = *(Oracle*)h;
#ifdef _OPENMP
  o.threads = threads > 0 ? threads : omp_get_max_threads();
#else
  (void)threads;
#endif
  return o.threads;
}

// Config presets, restated from the files that pin them (NOT read from libgsim: bench.py's reference
// arm must not map the product library).  LAN: [U] memberlist DefaultLANConfig as documented at
// agent/config/runtime.go:1271-1336; serf with Consul's overrides, internal/gossip/libserf/serf.go:19-36
// and agent/consul/config.go:622-623.  WAN: runtime.go:1348-1413.  Test harness: server_test.go:221-237.
void oracle_config_default_lan(gsim_config* c) {
  const uint64_t ms = 1000000ull, sec = 1000ull * ms, hour = 3600ull * sec;
  memset(c, 0, sizeof(*c));
  c->struct_size = (uint32_t)sizeof(*c);
  c->seed = 0x5EED0001ull;
  c->capacity = 1024;
  c->probe_interval_ns = sec;          // gossip_lan.probe_interval 1s
  c->probe_timeout_ns = 500 * ms;      // probe_timeout 500ms
  c->gossip_interval_ns = 200 * ms;    // gossip_interval 200ms
  c->gossip_nodes = 3;                 // gossip_nodes 3
  c->retransmit_mult = 4;              // retransmit_mult 4
  c->suspicion_mult = 4;               // suspicion_mult 4
  c->indirect_checks = 3;
  c->suspicion_max_timeout_mult = 6;
  c->awareness_max_multiplier = 8;
  c->gossip_to_the_dead_ns = 30 * sec;
  c->push_pull_interval_ns = 30 * sec;
  c->udp_buffer_size = 1400;
  c->event_buffer = 512;
  c->user_event_size_limit = 512;
  c->leave_propagate_delay_ns = 3 * sec;
  c->broadcast_timeout_ns = 5 * sec;
  c->reap_interval_ns = 15 * sec;
  c->reconnect_timeout_ns = 72 * hour;
  c->tombstone_timeout_ns = 24 * hour;
  c->world_size = 1;
  c->device = -1;
}
void oracle_config_default_wan(gsim_config* c) {
  const uint64_t ms = 1000000ull, sec = 1000ull * ms;
  oracle_config_default_lan(c);
  c->probe_interval_ns = 5 * sec;
  c->probe_timeout_ns = 3 * sec;
  c->gossip_interval_ns = 500 * ms;
  c->gossip_to_the_dead_ns = 60 * sec;
  c->push_pull_interval_ns = 60 * sec;
  c->suspicion_mult = 6;
}
void oracle_config_consul_test(gsim_config* c) {
  const uint64_t ms = 1000000ull;
  oracle_config_default_lan(c);
  c->probe_interval_ns = 100 * ms;
  c->probe_timeout_ns = 50 * ms;
  c->gossip_interval_ns = 100 * ms;
  c->suspicion_mult = 2;
}

int oracle_member_add(void* h, const gsim_member_desc* desc, uint32_t* id_out) {
  Oracle& o = *(Oracle*)h;
  if (!o.adjacency.empty()) return GSIM_ERR_STATE;  // static topology
  if (o.m.size() >= o.cfg.capacity) return GSIM_ERR_CAPACITY;
  int slot = free_slot(o);
  if (slot < 0) return GSIM_ERR_CAPACITY;
  uint32_t id = (uint32_t)o.m.size();
  o.m.emplace_back();
  o.pub.emplace_back();
  o.pub_change_tick.push_back(0);
  o.pub_coord.push_back(origin_coordinate());
  for (uint32_t b = 0; b < Oracle::RING; ++b) o.inbox[b].push_back(0);
  Member& me = o.m.back();
  me.v.truth = GSIM_TRUTH_UP;
  me.v.rank = GSIM_RANK_ALIVE;
  me.v.inc = 1;
  me.v.pending = 1;
  me.isolated = o.established > 0;  // with an empty base set there is nothing to be missing
  me.watched = desc && (desc->flags & GSIM_MEMBER_WATCHED);
  Stagger ph = stagger_of(o.cfg.seed, id, o.cfg.phase_group, o.P, o.GI);
  me.due = o.now + (ph.probe + o.P - o.now % o.P) % o.P;  // first tick >= now on its phase
  me.gossip_phase = ph.gossip;
  publish(o, id);
  o.up_count++;
  retune(o);
  if (desc && desc->name_len) o.name_lens.push_back(std::make_pair(id, desc->name_len));
  start_rumor(o, slot, GSIM_RUMOR_ALIVE, id, 1, 0, id,
              desc && desc->alive_msg_size ? desc->alive_msg_size : alive_bytes(o, id, 1, desc ? desc->meta_len : 0), 0);
  *id_out = id;
  return GSIM_OK;
}

int oracle_join(void* h, uint32_t id, const uint32_t* seeds, size_t n_seeds, int ignore_old, int* n_ok) {
  Oracle& o = *(Oracle*)h;
  if (id >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  if (o.m[id].v.truth != GSIM_TRUTH_UP) return GSIM_ERR_STATE;
  int ok = 0;
  for (size_t s = 0; s < n_seeds; ++s) {
    uint32_t sd = seeds[s];
    if (sd >= o.m.size() || sd == id || o.m[sd].v.truth != GSIM_TRUTH_UP) continue;
    merge_from(o, id, sd, ignore_old != 0);
    merge_from(o, sd, id, false);
    bool iso = o.m[id].isolated && o.m[sd].isolated;
    o.m[id].isolated = o.m[sd].isolated = iso;
    ++ok;
  }
  if (ok > 0) {
    uint32_t lt = o.m[id].ltime_member;  // [U] serf.broadcastJoin(clock.Time())
    int slot = free_slot(o);
    if (slot >= 0) start_rumor(o, slot, GSIM_RUMOR_JOIN_INTENT, id, 0, lt, id, intent_bytes(o, id, false, lt), 1);
    o.m[id].ltime_member = lt + 1;
  }
  if (n_ok) *n_ok = ok;
  return GSIM_OK;
}

int oracle_crash_many(void* h, const uint32_t* ids, size_t n) {
  Oracle& o = *(Oracle*)h;
  for (size_t x = 0; x < n; ++x) {
    if (ids[x] >= o.m.size()) return GSIM_ERR_NOT_FOUND;
    if (o.m[ids[x]].v.truth != GSIM_TRUTH_UP) continue;
    o.m[ids[x]].v.truth = GSIM_TRUTH_CRASHED;
    publish(o, ids[x]);
  }
  after_truth_change(o);
  return GSIM_OK;
}

int oracle_crash_fraction(void* h, uint32_t ppm, uint32_t salt, uint32_t* n_crashed) {
  Oracle& o = *(Oracle*)h;
  uint32_t thr = ppm >= 1000000 ? 0xFFFFFFFFu : (uint32_t)(((uint64_t)ppm << 32) / 1000000);
  uint32_t cnt = 0;
  for (uint32_t i = 0; i < o.m.size(); ++i) {
    if (o.m[i].v.truth != GSIM_TRUTH_UP) continue;
    if (philox4x32_10(o.cfg.seed, i, salt, PUR_CRASH, 0).v[0] >= thr) continue;
    o.m[i].v.truth = GSIM_TRUTH_CRASHED;
    publish(o, i);
    ++cnt;
  }
  if (n_crashed) *n_crashed = cnt;
  after_truth_change(o);
  return GSIM_OK;
}

int oracle_leave(void* h, uint32_t id) {
  Oracle& o = *(Oracle*)h;
  if (id >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  Member& me = o.m[id];
  if (me.v.truth != GSIM_TRUTH_UP || me.leaving) return GSIM_ERR_STATE;
  uint32_t lt = me.ltime_member;  // [U] serf.Leave: leave intent at clock.Time()
  int slot = free_slot(o);
  if (slot >= 0) start_rumor(o, slot, GSIM_RUMOR_LEAVE_INTENT, id, 0, lt, id, intent_bytes(o, id, true, lt), 1);
  Member& me2 = o.m[id];
  me2.ltime_member = lt + 1;
  me2.v.rank = GSIM_RANK_LEFT;  // [U] memberlist.Leave: dead{Node == From}
  me2.leaving = true;
  me2.change_tick = o.now;
  publish(o, id);
  if (o.cfg.flags & GSIM_FLAG_LOG_GLOBAL_EVENTS) host_event(o, GSIM_EVENT_MEMBER_LEAVE, id, NONE32, 0);
  uint32_t rounds = o.cfg.gossip_nodes ? (o.limit + o.cfg.gossip_nodes - 1) / o.cfg.gossip_nodes : 0;
  uint32_t drain = std::min(rounds * o.GI, to_ticks_ceil(o.cfg.broadcast_timeout_ns, o.tick_ns));
  uint32_t linger = 2 * drain + to_ticks_ceil(o.cfg.leave_propagate_delay_ns, o.tick_ns);
  o.shutdowns.push_back({o.now + linger, id});
  return GSIM_OK;
}

int oracle_force_leave(void* h, uint32_t via, uint32_t target, int prune) {
  Oracle& o = *(Oracle*)h;
  if (via >= o.m.size() || target >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  Member& me = o.m[target];
  if (me.v.rank == GSIM_RANK_DEAD) me.v.rank = GSIM_RANK_LEFT;  // [U] serf.RemoveFailedNode
  if (prune && me.v.rank == GSIM_RANK_LEFT && me.v.truth != GSIM_TRUTH_UP && me.v.truth != GSIM_TRUTH_NONE) {
    if (!me.v.pending) o.established--;
    me.v.truth = GSIM_TRUTH_NONE;
  }
  publish(o, target);
  after_truth_change(o);
  return GSIM_OK;
}


int oracle_user_event(void* h, uint32_t id, const void* name, size_t nl, const void* payload, size_t pl, int coalesce,
                      uint32_t* slot_out) {
  (void)coalesce;
  Oracle& o = *(Oracle*)h;
  if (id >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  if (nl + pl > o.cfg.user_event_size_limit) return GSIM_ERR_TOO_LARGE;
  if (o.m[id].v.truth != GSIM_TRUTH_UP) return GSIM_ERR_STATE;
  uint32_t lt = o.m[id].ltime_event;  // [U] serf.UserEvent: LTime = eventClock.Time(); Increment()
  std::string nm((const char*)name, nl), pd((const char*)payload, pl);
  for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r)
    if (((o.active >> r) & 1) && o.rumor[r].kind == GSIM_RUMOR_USER_EVENT && o.rumor[r].ltime == lt &&
        o.rumor[r].name == nm && o.rumor[r].payload == pd) {
      // the caller's own event buffer decides whether this is new to IT; the broadcast is queued anyway
      Member& me = o.m[id];
      if (!((me.heard >> r) & 1)) {
        me.heard |= 1u << r;
        o.rumor[r].heard_count++;
        if (o.rumor[r].heard_count == o.up_count && o.rumor[r].converged_tick == NONE32) o.rumor[r].converged_tick = o.now;
        if (me.watched) host_event(o, GSIM_EVENT_USER, r, id, lt);
      }
      me.queued |= 1u << r;
      me.tx[r] = 0;
      me.ltime_event = lt + 1;
      if (slot_out) *slot_out = r;
      return GSIM_OK;
    }
  uint32_t size = 1 + 1 + (6 + mp_uint(lt)) + (5 + mp_str(nl)) + (8 + mp_str(pl)) + (3 + 1);
  if (size > o.cfg.user_event_size_limit) return GSIM_ERR_TOO_LARGE;  // second check, on the encoded message
  int slot = free_slot(o);
  if (slot < 0) return GSIM_ERR_CAPACITY;
  start_rumor(o, slot, GSIM_RUMOR_USER_EVENT, id, 0, lt, id, size, 2);
  o.rumor[slot].name = nm;
  o.rumor[slot].payload = pd;
  o.m[id].ltime_event = lt + 1;
  if (o.m[id].watched) host_event(o, GSIM_EVENT_USER, (uint32_t)slot, id, lt);
  if (slot_out) *slot_out = (uint32_t)slot;
  return GSIM_OK;
}

// Out-of-band delivery of a tracked broadcast to one member (WAN bridge re-fire, config 5).
int oracle_rumor_inject(void* h, uint32_t slot, uint32_t id, int* accepted) {
  Oracle& o = *(Oracle*)h;
  if (accepted) *accepted = 0;
  if (slot >= GSIM_MAX_RUMORS || !((o.active >> slot) & 1)) return GSIM_ERR_NOT_FOUND;
  if (id >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  Member& me = o.m[id];
  if (me.v.truth != GSIM_TRUTH_UP) return GSIM_ERR_STATE;
  if ((me.heard >> slot) & 1) return GSIM_OK;
  Rumor& ru = o.rumor[slot];
  if (ru.kind == GSIM_RUMOR_USER_EVENT) {
    me.ltime_event = lamport_witness(me.ltime_event, ru.ltime);
    if (ru.ltime < me.event_min) return GSIM_OK;
    if (me.ltime_event > o.cfg.event_buffer && ru.ltime < me.ltime_event - o.cfg.event_buffer) return GSIM_OK;
    if (me.watched) host_event(o, GSIM_EVENT_USER, slot, id, ru.ltime);
  } else if (ru.kind == GSIM_RUMOR_JOIN_INTENT || ru.kind == GSIM_RUMOR_LEAVE_INTENT) {
    me.ltime_member = lamport_witness(me.ltime_member, ru.ltime);
  } else if (ru.kind == GSIM_RUMOR_ALIVE) {
    if (me.watched) host_event(o, GSIM_EVENT_MEMBER_JOIN, ru.subject, id, 0);
  } else if (ru.kind == GSIM_RUMOR_UPDATE) {
    if (me.watched) host_event(o, GSIM_EVENT_MEMBER_UPDATE, ru.subject, id, 0);
  }
  me.heard |= 1u << slot;
  me.queued |= 1u << slot;
  me.tx[slot] = 0;
  ru.heard_count++;
  if (ru.heard_count == o.up_count && ru.converged_tick == NONE32) ru.converged_tick = o.now;
  if (accepted) *accepted = 1;
  return GSIM_OK;
}

// (*Serf).SetTags -> [U] memberlist.UpdateNode: alive{nextIncarnation(), new meta} is broadcast.
int oracle_member_update(void* h, uint32_t id, uint32_t alive_msg_size, uint32_t* slot_out) {
  Oracle& o = *(Oracle*)h;
  if (id >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  if (o.m[id].v.truth != GSIM_TRUTH_UP || o.m[id].leaving) return GSIM_ERR_STATE;
  int slot = free_slot(o);
  if (slot < 0) return GSIM_ERR_CAPACITY;
  Member& me = o.m[id];
  me.v.inc += 1;
  publish(o, id);
  start_rumor(o, slot, GSIM_RUMOR_UPDATE, id, me.v.inc, 0, id, alive_msg_size ? alive_msg_size : alive_bytes(o, id, me.v.inc, 0), 0);
  if (slot_out) *slot_out = (uint32_t)slot;
  return GSIM_OK;
}

// CSR peer graph: adjacency[i] = col_idx[row_ptr[i] .. row_ptr[i+1])
int oracle_graph_set(void* h, uint32_t n_rows, const uint32_t* row_ptr, const uint32_t* col_idx) {
  Oracle& o = *(Oracle*)h;
  o.adjacency.clear();
  if (!n_rows) return GSIM_OK;
  if (n_rows != o.m.size()) return GSIM_ERR_INVALID;
  o.adjacency.resize(n_rows);
  for (uint32_t i = 0; i < n_rows; ++i) o.adjacency[i].assign(col_idx + row_ptr[i], col_idx + row_ptr[i + 1]);
  return GSIM_OK;
}

int oracle_member_reconnect_timeout_set(void* h, uint32_t id, uint64_t timeout_ns) {
  Oracle& o = *(Oracle*)h;
  if (id >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  o.m[id].own_reconnect_timeout_ns = timeout_ns;
  return GSIM_OK;
}

int oracle_coordinate_get(void* h, uint32_t id, double out[11]) {
  Oracle& o = *(Oracle*)h;
  if (!(o.cfg.flags & GSIM_FLAG_COORDINATES)) return GSIM_ERR_STATE;
  if (id >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  const Coordinate& c = o.m[id].coord;
  for (int k = 0; k < 8; ++k) out[k] = c.v[k];
  out[8] = c.err;
  out[9] = c.adj;
  out[10] = c.h;
  return GSIM_OK;
}

int oracle_member_watch(void* h, uint32_t id, int on) {
  Oracle& o = *(Oracle*)h;
  if (id >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  o.m[id].watched = on != 0;
  return GSIM_OK;
}

int oracle_latency_set(void* h, uint32_t n_dcs, const uint8_t* lat) {
  Oracle& o = *(Oracle*)h;
  if (n_dcs > 64) return GSIM_ERR_INVALID;
  const uint32_t depth = o.cfg.mailbox_depth ? o.cfg.mailbox_depth : 2;
  for (uint32_t x = 0; x < n_dcs * n_dcs; ++x)
    if (lat[x] < 1 || lat[x] >= depth) return GSIM_ERR_INVALID;
  for (uint32_t a = 0; a < n_dcs; ++a)
    for (uint32_t b = 0; b < n_dcs; ++b) o.latency[a][b] = lat[a * n_dcs + b];
  o.n_dcs = n_dcs;
  return GSIM_OK;
}

int oracle_step(void* h, uint32_t ticks) {
  step(*(Oracle*)h, ticks);
  return GSIM_OK;
}
uint32_t oracle_now(void* h) { return ((Oracle*)h)->now; }

int oracle_run_until(void* h, int predicate, uint32_t arg, uint32_t max_ticks, uint32_t check_every, uint32_t* tick_out) {
  Oracle& o = *(Oracle*)h;
  if (!check_every) return GSIM_ERR_INVALID;
  if (tick_out) *tick_out = NONE32;
  uint32_t done = 0;
  for (;;) {
    uint32_t result = NONE32;
    if (predicate == GSIM_PRED_RUMOR_CONVERGED) {
      if (arg >= GSIM_MAX_RUMORS) return GSIM_ERR_INVALID;
      result = ((o.active >> arg) & 1) ? o.rumor[arg].converged_tick : NONE32;
    } else if (predicate == GSIM_PRED_ALL_RUMORS_CONVERGED) {
      bool all = true;
      uint32_t mx = 0;
      for (int r = 0; r < GSIM_MAX_RUMORS; ++r)
        if ((o.active >> r) & 1) {
          if (o.rumor[r].converged_tick == NONE32) all = false;
          else mx = std::max(mx, o.rumor[r].converged_tick);
        }
      if (all) result = mx;
    } else if (predicate == GSIM_PRED_CRASHED_ALL_DEAD) {
      result = o.crashed_dead_tick;
    } else {
      return GSIM_ERR_INVALID;
    }
    if (result != NONE32) {
      if (tick_out) *tick_out = result;
      break;
    }
    if (done >= max_ticks) break;
    uint32_t chunk = std::min(check_every, max_ticks - done);
    step(o, chunk);
    done += chunk;
  }
  return GSIM_OK;
}

int oracle_members(void* h, uint32_t observer, gsim_member* out, size_t cap, size_t* n) {
  Oracle& o = *(Oracle*)h;
  if (observer >= o.m.size()) return GSIM_ERR_NOT_FOUND;
  size_t cnt = 0;
  for (uint32_t c = 0; c < o.m.size(); ++c) {
    const View& v = o.m[c].v;
    if (v.truth == GSIM_TRUTH_NONE) continue;
    if (!o.adjacency.empty() && c != observer &&
        std::find(o.adjacency[observer].begin(), o.adjacency[observer].end(), c) == o.adjacency[observer].end())
      continue;
    if (!knows(o, observer, o.m[observer], c)) continue;
    if (out && cnt < cap) {
      out[cnt].id = c;
      out[cnt].incarnation = v.inc;
      out[cnt].rank = v.rank;
      out[cnt].status = v.rank == GSIM_RANK_DEAD ? GSIM_STATUS_FAILED : v.rank == GSIM_RANK_LEFT ? GSIM_STATUS_LEFT : GSIM_STATUS_ALIVE;
    }
    ++cnt;
  }
  if (n) *n = cnt;
  return GSIM_OK;
}

int oracle_poll_events(void* h, gsim_event* out, size_t cap, size_t* n) {
  Oracle& o = *(Oracle*)h;
  std::sort(o.events.begin(), o.events.end(), [](const gsim_event& a, const gsim_event& b) {
    if (a.tick != b.tick) return a.tick < b.tick;
    if (a.type != b.type) return a.type < b.type;
    if (a.subject != b.subject) return a.subject < b.subject;
    return a.observer < b.observer;
  });
  size_t take = std::min(cap, o.events.size());
  for (size_t x = 0; x < take; ++x) out[x] = o.events[x];
  o.events.erase(o.events.begin(), o.events.begin() + take);
  *n = take;
  return GSIM_OK;
}

int oracle_rumor_info_get(void* h, uint32_t slot, gsim_rumor_info* out) {
  Oracle& o = *(Oracle*)h;
  if (slot >= GSIM_MAX_RUMORS || !((o.active >> slot) & 1)) return GSIM_ERR_NOT_FOUND;
  Counts c = count_all(o);
  const Rumor& ru = o.rumor[slot];
  out->kind = ru.kind;
  out->subject = ru.subject;
  out->incarnation = ru.inc;
  out->ltime = ru.ltime;
  out->origin = ru.origin;
  out->size_bytes = ru.size;
  out->start_tick = ru.start;
  out->heard_count = c.heard[slot];
  out->queued_count = c.queued[slot];
  out->converged_tick = ru.converged_tick;
  return GSIM_OK;
}

int oracle_rumor_retire(void* h, uint32_t slot) {
  Oracle& o = *(Oracle*)h;
  if (slot >= GSIM_MAX_RUMORS || !((o.active >> slot) & 1)) return GSIM_ERR_NOT_FOUND;
  if (o.rumor[slot].kind == GSIM_RUMOR_ALIVE) {
    Counts c = count_all(o);
    if (c.heard[slot] != o.up_count || c.isolated_up) return GSIM_ERR_STATE;
  }
  retire(o, slot);
  return GSIM_OK;
}

int oracle_stats_get(void* h, gsim_stats* out) {
  Oracle& o = *(Oracle*)h;
  memset(out, 0, sizeof(*out));
  memcpy(out->counters, o.stats, sizeof(o.stats));
  Counts c = count_all(o);
  out->node_ticks = o.node_ticks;
  out->tick = o.now;
  out->n_members = (uint32_t)o.m.size();
  out->n_up = c.truth[GSIM_TRUTH_UP];
  out->n_crashed = c.truth[GSIM_TRUTH_CRASHED];
  out->n_gone = c.truth[GSIM_TRUTH_GONE];
  out->n_view_alive = c.rank[0];
  out->n_view_suspect = c.rank[1];
  out->n_view_dead = c.rank[2];
  out->n_view_left = c.rank[3];
  out->retransmit_limit = o.limit;
  out->suspicion_k = o.sus_k;
  for (int q = 0; q < MAX_SUS; ++q) out->suspicion_ticks[q] = o.sus_ticks[q];
  out->probe_interval_ticks = o.P;
  out->probe_timeout_ticks = o.T;
  out->gossip_interval_ticks = o.GI;
  out->events_dropped = o.events_dropped;
  return GSIM_OK;
}

// Same canonical digest as gsim_state_hash: live fields only, summed over rows.
int oracle_state_hash(void* h, uint64_t out[4]) {
  Oracle& o = *(Oracle*)h;
  uint64_t lanes[4] = {0, 0, 0, 0};
  auto fold = [&](uint64_t x) {
    lanes[0] += x;
    lanes[1] += mix(x, 1);
    lanes[2] += mix(x, 2);
    lanes[3] += mix(x, 3);
  };
  // pending accusations per subject, as the mailbox would hold them
  std::vector<Accusation> acc = o.arriving;
  std::sort(acc.begin(), acc.end(), [](const Accusation& a, const Accusation& b) {
    if (a.subject != b.subject) return a.subject < b.subject;
    if (a.inc != b.inc) return a.inc > b.inc;
    return a.from < b.from;
  });
  // pending push-pull records per receiver
  std::vector<PushPull> pp = o.pp_arriving;
  std::sort(pp.begin(), pp.end(), [](const PushPull& a, const PushPull& b) {
    if (a.to != b.to) return a.to < b.to;
    return a.from < b.from;
  });
  size_t ap = 0, pq = 0;
  for (uint32_t i = 0; i < o.m.size(); ++i) {
    const Member& me = o.m[i];
    while (ap < acc.size() && acc[ap].subject < i) ++ap;
    while (pq < pp.size() && pp[pq].to < i) ++pq;
    if (me.v.truth == GSIM_TRUTH_NONE) continue;
    const bool up = me.v.truth == GSIM_TRUTH_UP;
    const bool probing = up && me.stage != ST_IDLE;
    const bool sus = me.v.rank == GSIM_RANK_SUSPECT;
    uint64_t x = 0x9E3779B97F4A7C15ull;
    x = mix(x, i);
    x = mix(x, pack_key(me.v));
    x = mix(x, pack_meta(me));
    x = mix(x, up ? me.due : 0);
    x = mix(x, me.cursor);
    x = mix(x, me.pass);
    x = mix(x, probing ? me.probe_target : 0);
    x = mix(x, probing ? me.probe_inc : 0);
    x = mix(x, sus ? me.sus_start : 0);
    for (int q = 0; q < MAX_SUS; ++q) x = mix(x, sus ? me.sus_from[q] : 0);
    x = mix(x, me.v.rank >= GSIM_RANK_DEAD ? me.change_tick : 0);
    x = mix(x, me.ltime_member);
    x = mix(x, me.ltime_event);
    x = mix(x, me.event_min);
    uint32_t heard = me.heard & o.active;
    x = mix(x, heard);
    x = mix(x, me.queued & o.active);
    const bool has_pp = pq < pp.size() && pp[pq].to == i;
    bool has_acc = (ap < acc.size() && acc[ap].subject == i) || has_pp;  // bit 31: auxiliary mail
    uint32_t inb = (o.inbox[o.now % Oracle::RING][i] & o.active) | (has_acc ? 0x80000000u : 0);
    x = mix(x, inb);
    // packets still in flight on a WAN pool, nearest arrival first (what a mailbox ring of
    // cfg.mailbox_depth slots can hold beyond the slot read next and the one just emptied)
    for (uint32_t ahead = 1; ahead + 1 < (o.cfg.mailbox_depth ? o.cfg.mailbox_depth : 2); ++ahead)
      x = mix(x, o.inbox[(o.now + ahead) % Oracle::RING][i] & o.active);
    for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r)
      if ((heard >> r) & 1) x = mix(x, (r << 8) | me.tx[r]);
    if (o.cfg.flags & GSIM_FLAG_COORDINATES) {
      const Coordinate& c = me.coord;
      const double words[11] = {c.v[0], c.v[1], c.v[2], c.v[3], c.v[4], c.v[5], c.v[6], c.v[7], c.err, c.adj, c.h};
      for (double w : words) {
        uint64_t bits;
        memcpy(&bits, &w, 8);
        x = mix(x, bits);
      }
      x = mix(x, me.adj_index);
    }
    if (has_acc) {
      int cnt = 0;
      size_t q = ap;
      for (; q < acc.size() && acc[q].subject == i && cnt < MAX_SUS; ++q, ++cnt)
        x = mix(x, ((uint64_t)(~acc[q].inc) << 32) | acc[q].from);
      for (; cnt < MAX_SUS; ++cnt) x = mix(x, 0xFFFFFFFFFFFFFFFFull);
      if (o.pp_every) {  // who is waiting for an answer (smallest ids), and the partners' clocks
        uint32_t waiting = 0, cm = 0, ce = 0;
        for (size_t y = pq; y < pp.size() && pp[y].to == i; ++y) {
          if (pp[y].request && waiting < PUSHPULL_BACKLOG) {
            x = mix(x, pp[y].from);
            ++waiting;
          }
          cm = std::max(cm, pp[y].clock_member);
          ce = std::max(ce, pp[y].clock_event);
        }
        for (; waiting < PUSHPULL_BACKLOG; ++waiting) x = mix(x, 0xFFFFFFFFu);
        x = mix(x, cm);
        x = mix(x, ce);
      }
    }
    fold(x);
  }
  uint64_t g = mix(0x243F6A8885A308D3ull, o.now);
  g = mix(g, o.m.size());
  g = mix(g, o.up_count);
  g = mix(g, o.active);
  for (uint32_t r = 0; r < GSIM_MAX_RUMORS; ++r)
    if ((o.active >> r) & 1) {
      g = mix(g, ((uint64_t)r << 32) | o.rumor[r].kind);
      g = mix(g, ((uint64_t)o.rumor[r].subject << 32) | o.rumor[r].ltime);
    }
  fold(g);
  memcpy(out, lanes, sizeof(lanes));
  return GSIM_OK;
}

// Columns in the layout of gsim_column_read so tests can compare arrays element-wise.
int oracle_column_read(void* h, int column, void* out, size_t cap_bytes, size_t* n_bytes) {
  Oracle& o = *(Oracle*)h;
  const size_t cap = o.cfg.capacity, n = o.m.size();
  size_t bytes = cap * 4;
  if (column == GSIM_COL_SUS_FROM) bytes = cap * 4 * MAX_SUS;
  if (column == GSIM_COL_TX) bytes = cap * GSIM_MAX_RUMORS;
  if (n_bytes) *n_bytes = bytes;
  if (cap_bytes < bytes) return GSIM_ERR_INVALID;
  memset(out, 0, bytes);
  uint32_t* w = (uint32_t*)out;
  uint8_t* b = (uint8_t*)out;
  if (column == GSIM_COL_SUS_FROM) memset(out, 0xFF, bytes);
  for (size_t i = 0; i < n; ++i) {
    const Member& me = o.m[i];
    switch (column) {
      case GSIM_COL_KEY: w[i] = pack_key(me.v); break;
      case GSIM_COL_META: w[i] = pack_meta(me); break;
      case GSIM_COL_DUE: w[i] = me.due; break;
      case GSIM_COL_CURSOR: w[i] = me.cursor; break;
      case GSIM_COL_PASS: w[i] = me.pass; break;
      case GSIM_COL_PROBE_TGT: w[i] = me.probe_target; break;
      case GSIM_COL_PROBE_INC: w[i] = me.probe_inc; break;
      case GSIM_COL_SUS_START: w[i] = me.sus_start; break;
      case GSIM_COL_SUS_FROM:
        for (int q = 0; q < MAX_SUS; ++q) w[(size_t)q * cap + i] = me.sus_from[q];
        break;
      case GSIM_COL_CHANGE_TICK: w[i] = me.change_tick; break;
      case GSIM_COL_LTIME_MEMBER: w[i] = me.ltime_member; break;
      case GSIM_COL_LTIME_EVENT: w[i] = me.ltime_event; break;
      case GSIM_COL_EVENT_MIN: w[i] = me.event_min; break;
      case GSIM_COL_HEARD: w[i] = me.heard; break;
      case GSIM_COL_QUEUED: w[i] = me.queued; break;
      case GSIM_COL_TX:
        for (int r = 0; r < GSIM_MAX_RUMORS; ++r) b[(size_t)r * cap + i] = me.tx[r];
        break;
      case GSIM_COL_INBOX: w[i] = o.inbox[o.now % Oracle::RING][i] & 0x7FFFFFFFu; break;
      default: return GSIM_ERR_INVALID;
    }
  }
  if (column == GSIM_COL_INBOX) {
    for (const Accusation& a : o.arriving) w[a.subject] |= 0x80000000u;
    for (const PushPull& a : o.pp_arriving) w[a.to] |= 0x80000000u;
  }
  return GSIM_OK;
}

}  // extern "C"
