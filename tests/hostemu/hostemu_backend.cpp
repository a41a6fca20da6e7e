// hostemu_backend.cpp — TEST INFRASTRUCTURE ONLY.  Never linked into libgsim.so and never
// loaded by the consul_b200 package.
//
// The tick kernel's per-row body (consul_b200/csrc/gs_row.h) is plain C++ between the
// atomics macros, so it can be compiled by g++ and looped on the CPU.  That lets the
// GPU-less development container run the kernel LOGIC against the oracle before any GPU
// time is spent, and lets the tests execute rows in adversarial orders (reverse, strided)
// to check that results do not depend on scheduling — the property that makes the
// CUDA launch deterministic.  It is built into tests/hostemu/libgsim_hostemu.so together
// with gs_api.cpp by `__graft_entry__.build()`; parity claims are made only for the CUDA
// backend on a real B200 (tests marked `gpu`).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <vector>

#include "../../consul_b200/csrc/gs_aux.h"
#include "../../consul_b200/csrc/gs_backend.h"

namespace {

struct HostSink {
  static constexpr bool kCoords = true;  // the host build always carries the coordinate update
  uint64_t* stats;
  uint32_t* heard_cnt;
  uint32_t local_heard[32];
  bool active = false;           // this launch saw mail or posted some
  uint32_t min_horizon = GS_NEVER;
  void activity() { active = true; }
  void horizon(uint32_t h) {
    if (h < min_horizon) min_horizon = h;
  }
  // what gs_q_publish does at the end of a launch: every rank's copy of the scheduling words
  void publish(const GsDev& d, const GsGlobals& g, uint32_t t, bool window) {
    for (uint32_t r = 0; r < (g.world ? g.world : 1u); ++r) {
      uint32_t* qs = d.qstate[r];
      if (active && !window) {
        uint32_t old = __atomic_load_n(qs + GS_Q_LAST_ACTIVE, __ATOMIC_RELAXED);
        while (old < t + 1u && !__atomic_compare_exchange_n(qs + GS_Q_LAST_ACTIVE, &old, t + 1u, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
        }
      }
      if (min_horizon != GS_NEVER) {
        uint32_t old = __atomic_load_n(qs + GS_Q_HORIZON, __ATOMIC_RELAXED);
        while (old > min_horizon && !__atomic_compare_exchange_n(qs + GS_Q_HORIZON, &old, min_horizon, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
        }
      }
    }
    if (active && window) d.qstate[g.rank][GS_Q_VIOLATION] = t + 1u;
  }
  void stat(int idx, uint32_t v) { __atomic_fetch_add(&stats[idx], (uint64_t)v, __ATOMIC_RELAXED); }
  void heard(uint32_t r) { local_heard[r] += 1; }
  void crashed_dead(const GsDev& d, uint32_t t) {
    uint32_t old = __atomic_fetch_sub(d.crashed_alive, 1u, __ATOMIC_RELAXED);
    if (old == 1u) *d.crashed_dead_tick = t;
  }
  void log_event(const GsDev& d, const GsGlobals& g, uint32_t t, uint32_t type, uint32_t subject,
                 uint32_t observer, uint32_t ltime) {
    uint32_t pos = __atomic_fetch_add(&d.evlog_cursor[0], 1u, __ATOMIC_RELAXED);
    if (pos < g.evlog_cap) {
      GsEventRec e = {t, type, subject, observer, ltime, 0u};
      d.evlog[pos] = e;
    } else {
      d.evlog_cursor[1]++;
    }
  }
};

class HostEmuBackend : public GsBackend {
 public:
  HostEmuBackend() {
    const char* o = getenv("GSIM_HOSTEMU_ORDER");
    order_ = o ? atoi(o) : 0;
    no_fast_ = getenv("GSIM_HOSTEMU_NO_FAST") != nullptr;  // generic path only (debugging aid)
    err_[0] = 0;
  }
  const char* name() const override { return "hostemu (tests only)"; }
  void* alloc(size_t bytes) override { return calloc(1, bytes ? bytes : 4); }
  void release(void* p) override {
    if (!sharded_) free(p);
  }
  bool h2d(void* dst, const void* src, size_t bytes) override {
    memcpy(dst, src, bytes);
    return true;
  }
  bool d2h(void* dst, const void* src, size_t bytes) override {
    memcpy(dst, src, bytes);
    return true;
  }
  bool row_read(const GsDev& d, uint32_t i, uint32_t out[8]) override {
    const uint32_t* col[8] = {d.key[0], d.key[1], d.meta, d.heard, d.queued, d.ltime_member, d.ltime_event, d.event_min};
    for (int x = 0; x < 8; ++x) out[x] = col[x][i];
    return true;
  }
  bool fill32(uint32_t* dst, uint32_t value, size_t count) override {
    for (size_t i = 0; i < count; ++i) dst[i] = value;
    return true;
  }
  bool fill8(uint8_t* dst, uint8_t value, size_t count) override {
    memset(dst, value, count);
    return true;
  }
  bool init_rows(const GsDev& d, const GsGlobals*, const GsGlobals& g, uint32_t first,
                 uint32_t count, uint32_t now) override {
    for (uint32_t x = 0; x < count; ++x) gs_init_row(d, g, first + x, now);
    return true;
  }
  uint32_t row_at(uint32_t x, uint32_t n) const {
    if (order_ == 1) return n - 1 - x;  // reverse
    if (order_ == 2) {                  // odd rows first, then even rows
      uint32_t odd = n / 2;
      return x < odd ? 2 * x + 1 : 2 * (x - odd);
    }
    return x;
  }
  bool run_ticks(const GsDev& d, const GsGlobals* g_dev, const GsGlobals&, uint32_t t0,
                 uint32_t nticks, bool, double*, uint64_t* launches, const GsXbar* xbar) override {
    const GsGlobals& g = *g_dev;  // the kernels read the device copy
    for (uint32_t k = 0; k < nticks; ++k) {
      const uint32_t t = *d.tick_base + k;
      if (t != t0 + k) {
        snprintf(err_, sizeof(err_), "tick_base out of sync");
        return false;
      }
      const uint32_t gslot = t % g.GI;
      HostSink sink;
      sink.stats = reinterpret_cast<uint64_t*>(d.stats);
      sink.heard_cnt = d.heard_cnt;
      memset(sink.local_heard, 0, sizeof(sink.local_heard));
      sink.active = false;
      sink.min_horizon = GS_NEVER;
      // same activity test as gs_tick_kernel: mailbox word, plus `due` only for tiles whose
      // ticker phase can be due at this tick
      const uint32_t pslot = t % g.P;
      // this rank's members: everything, or its contiguous range on a sharded pool
      uint32_t lo = 0, hi = g.n;
      if (g.world > 1u) {
        lo = g.rank * g.rows_per_rank < g.n ? g.rank * g.rows_per_rank : g.n;
        hi = lo + g.rows_per_rank < g.n ? lo + g.rows_per_rank : g.n;
      }
      for (uint32_t x = 0; x < hi - lo; ++x) {
        const uint32_t i = lo + row_at(x, hi - lo);
        const uint32_t inb = d.inbox[t & g.ring_mask][i];
        uint32_t due = GS_NEVER;
        if (gs_tile_probe_gate(g, i / GS_TILE, pslot, (t + g.P - g.T % g.P) % g.P)) due = d.due[i];
        if (!(inb != 0u || due == t || gs_pp_due(g.pp_interval, g.rot_pp, i / g.phase_group, t))) continue;
        if (inb == 0u && due == t && !no_fast_) {  // same two tiers as the kernel
          GsFastProbe f;
          bool acked = false;
          gs_fast_load(d, t & 1u, i, f);
          if (gs_fast_target(d, g, t & 1u, i, f) && gs_fast_finish(d, g, sink, i, t, f, &acked)) {
            sink.stat(GS_ST_PROBES, 1);
            sink.stat(GS_ST_ACTIVE_ROWS, 1);
            if (acked) sink.stat(GS_ST_ACKS, 1);
            continue;
          }
        }
        gs_row_step(d, g, i, t, gslot, inb, sink);
      }
      for (uint32_t r = 0; r < GS_MAX_RUMORS; ++r) {
        uint32_t c = sink.local_heard[r];
        if (c) {
          uint32_t old = __atomic_fetch_add(&d.heard_cnt[r], c, __ATOMIC_RELAXED);
          if (old + c == g.up_count) d.conv_tick[r] = t;
        }
      }
      sink.publish(d, g, t, false);
      ++launches_;
      if (xbar) xbar_host(*xbar);
    }
    *d.tick_base += nticks;
    ++launches_;
    if (launches) *launches += nticks;
    return true;
  }
  // gs_window_kernel on the host.  Rows are taken one after the other and each runs ALL its ticks of
  // the window before the next row is looked at — as far from lock-step as an order can be, which is
  // the point: inside a quiet window rows are independent.
  bool run_windows(const GsDev& d, const GsGlobals* g_dev, const GsGlobals&, uint32_t t0, uint32_t nticks,
                   uint32_t per_launch, bool, double*, uint64_t* launches, uint32_t* ticks_done,
                   const GsXbar* xbar, bool pristine) override {
    const GsGlobals& g = *g_dev;
    *ticks_done = 0;
    if (!nticks || !g.n) return true;
    uint32_t* qs = d.qstate[g.rank];
    qs[GS_Q_WIN_END] = t0;
    qs[GS_Q_VIOLATION] = 0u;
    if (*d.tick_base != t0) {
      snprintf(err_, sizeof(err_), "tick_base out of sync");
      return false;
    }
    const uint32_t K = per_launch < g.P ? g.P : per_launch;
    uint32_t lo = 0, hi = g.n;
    if (g.world > 1u) {
      lo = g.rank * g.rows_per_rank < g.n ? g.rank * g.rows_per_rank : g.n;
      hi = lo + g.rows_per_rank < g.n ? lo + g.rows_per_rank : g.n;
    }
    uint32_t n_launch = 0;
    for (uint32_t w0 = t0; w0 < t0 + nticks; w0 += K) {
      const uint32_t n_ticks = t0 + nticks - w0 < K ? t0 + nticks - w0 : K;
      ++n_launch;
      ++launches_;
      const uint32_t reached = __atomic_load_n(qs + GS_Q_WIN_END, __ATOMIC_RELAXED);
      const uint32_t horizon = __atomic_load_n(qs + GS_Q_HORIZON, __ATOMIC_RELAXED);
      if (reached < w0) continue;
      uint32_t w1 = w0 + n_ticks;
      if (horizon < w1) w1 = horizon;
      if (w1 <= w0) continue;
      HostSink sink;
      sink.stats = reinterpret_cast<uint64_t*>(d.stats);
      sink.heard_cnt = d.heard_cnt;
      memset(sink.local_heard, 0, sizeof(sink.local_heard));
      uint32_t spec[GS_MAX_SPECIAL];
      const uint32_t n_spec = gs_special_members(g, spec);
      const bool closed = pristine && !no_fast_ && g.loss_thr == 0u && g.graph_n == 0u && d.coord == nullptr &&
                          g.pp_interval == 0u && n_spec <= GS_MAX_SPECIAL;
      for (uint32_t x = 0; x < hi - lo; ++x) {
        const uint32_t i = lo + row_at(x, hi - lo);
        const uint32_t pp = gs_probe_phase(g.rot_p, (i / GS_TILE) >> g.phase_shift, g.P);
        // every tick of the launch at which this member can be due: congruent to its ticker phase or to
        // phase + ProbeTimeout (a launch covers one ProbeInterval in general, many when no probe can fail)
        for (uint32_t t = w0; t < w1; ++t) {
          if (t % g.P != pp && t % g.P != (pp + g.T) % g.P) continue;
          if (d.due[i] != t) continue;
          if (closed && t % g.P == pp) {
            // the closed form of the window kernel (gs_pristine_probes), under the kernel's own conditions:
            // the ticker fires, the member is up, listed alive, idle
            const uint32_t k0 = d.key[t & 1u][i], m = d.meta[i];
            if (gs_key_truth(k0) == GS_TRUTH_UP && gs_key_rank(k0) == GS_RANK_ALIVE && gs_meta_stage(m) == GS_STAGE_IDLE &&
                !(m & (GS_META_DIRTY | GS_META_ISOLATED))) {
              const GsU4 rk = gs_perm_keys(g.seed_lo, g.seed_hi, i, d.pass[i]);
              const uint32_t k = gs_pristine_probes(g.n, g.perm_bits, rk, i, d.cursor[i], t, w1, g.P, spec, n_spec);
              if (k) {
                const uint32_t aw = gs_meta_aw(m);
                d.meta[i] = gs_meta_set_aw(m, aw > k ? aw - k : 0u);
                d.due[i] = t + k * g.P;
                d.cursor[i] += k;
                sink.stat(GS_ST_PROBES, k);
                sink.stat(GS_ST_ACTIVE_ROWS, k);
                sink.stat(GS_ST_ACKS, k);
                continue;
              }
            }
          }
          if (!no_fast_) {
            GsFastProbe f;
            bool acked = false;
            gs_fast_load(d, t & 1u, i, f);
            if (gs_fast_target(d, g, t & 1u, i, f) && gs_fast_finish(d, g, sink, i, t, f, &acked)) {
              sink.stat(GS_ST_PROBES, 1);
              sink.stat(GS_ST_ACTIVE_ROWS, 1);
              if (acked) sink.stat(GS_ST_ACKS, 1);
              continue;
            }
          }
          gs_row_step(d, g, i, t, t % g.GI, 0u, sink);
        }
      }
      if (K > g.P && sink.min_horizon != GS_NEVER) sink.active = true;  // a probe went unanswered in a long launch
      sink.publish(d, g, w0, true);
      qs[GS_Q_WIN_END] = w1;
      if (xbar) xbar_host(*xbar);  // sharded: one inter-rank barrier per window
    }
    *d.tick_base = qs[GS_Q_WIN_END];
    ++launches_;
    if (launches) *launches += n_launch;
    if (qs[GS_Q_VIOLATION] != 0u) {
      snprintf(err_, sizeof(err_), "quiet window starting at tick %u met mail or posted some", qs[GS_Q_VIOLATION] - 1u);
      return false;
    }
    *ticks_done = qs[GS_Q_WIN_END] - t0;
    return true;
  }
  bool quiet_scan(const GsDev& d, const GsGlobals* g_dev, const GsGlobals&, uint32_t now, uint32_t first,
                  uint32_t count) override {
    const GsGlobals& g = *g_dev;
    HostSink sink;
    for (uint32_t x = 0; x < count; ++x) {
      const uint32_t i = first + x;
      if (gs_key_truth(d.key[now & 1u][i]) != GS_TRUTH_UP) continue;
      const uint32_t stage = gs_meta_stage(d.meta[i]);
      if (stage == GS_STAGE_IDLE) continue;
      sink.horizon(stage == GS_STAGE_WAIT_T ? d.due[i] - g.T + g.P : d.due[i]);
    }
    sink.publish(d, g, now, false);
    ++launches_;
    return true;
  }
  bool crash_fraction(const GsDev& d, const GsGlobals*, const GsGlobals& g, uint32_t thr,
                      uint32_t salt, uint32_t, uint32_t* n_crashed) override {
    uint32_t c = 0;
    for (uint32_t i = 0; i < g.n; ++i) c += gs_crash_row(d, g, i, thr, salt) ? 1u : 0u;
    *n_crashed = c;
    return true;
  }
  bool reap_rows(const GsDev& d, const GsGlobals*, const GsGlobals& g, uint32_t now,
                 uint32_t reconnect_ticks, uint32_t tombstone_ticks, bool log_events,
                 uint32_t counts[2]) override {
    counts[0] = counts[1] = 0;
    HostSink sink;
    memset(&sink, 0, sizeof(sink));
    for (uint32_t i = 0; i < g.n; ++i) {
      const uint32_t r = gs_reap_row(d, g, i, now, reconnect_ticks, tombstone_ticks);
      counts[0] += r & 1u;
      counts[1] += (r >> 1) & 1u;
      if (r && log_events) sink.log_event(d, g, now, GS_EV_MEMBER_REAP, i, GS_EMPTY32, 0u);
    }
    return true;
  }
  bool recount(const GsDev& d, const GsGlobals*, const GsGlobals& g, uint32_t now, uint32_t first, uint32_t count,
               GsRecount* out) override {
    memset(out, 0, sizeof(*out));
    for (uint32_t i = first; i < g.n && i - first < count; ++i) {
      uint32_t k = d.key[now & 1u][i];
      if (gs_key_truth(k) != GS_TRUTH_NONE && gs_key_pending(k)) out->pending++;
      uint32_t truth = gs_key_truth(k), rank = gs_key_rank(k);
      out->truth_cnt[truth]++;
      if (truth != GS_TRUTH_NONE) out->rank_cnt[rank]++;
      if (truth == GS_TRUTH_CRASHED && rank < GS_RANK_DEAD) out->crashed_alive++;
      if ((truth == GS_TRUTH_CRASHED || truth == GS_TRUTH_GONE) && rank < GS_RANK_DEAD) out->unreachable_live++;
      if (truth == GS_TRUTH_UP && (d.meta[i] & GS_META_ISOLATED)) out->isolated_up++;
      if (truth == GS_TRUTH_UP) {
        uint32_t h = d.heard[i] & g.active_mask, q = d.queued[i] & g.active_mask;
        for (uint32_t r = 0; r < GS_MAX_RUMORS; ++r) {
          out->heard_cnt[r] += (h >> r) & 1u;
          out->queued_cnt[r] += (q >> r) & 1u;
        }
      }
    }
    return true;
  }
  bool state_hash(const GsDev& d, const GsGlobals*, const GsGlobals& g, uint32_t now,
                  uint64_t out[4]) override {
    out[0] = out[1] = out[2] = out[3] = 0;
    for (uint32_t i = 0; i < g.n; ++i) {
      uint64_t h = gs_hash_row(d, g, i, now);
      if (!h) continue;
      uint64_t lanes[4];
      gs_hash_lanes(h, lanes);
      for (int q = 0; q < 4; ++q) out[q] += lanes[q];
    }
    return true;
  }
  // ---- sharded pools on the host: the same layout as gs_vmm.h with memfd + mmap(MAP_FIXED) ----
  // standing in for cuMemCreate + cuMemMap, so the controller protocol, the descriptor
  // exchange and the per-tick barrier can be tested with 2 gloo processes on a GPU-less box.
  struct Col {
    uint8_t* va;
    size_t slice_bytes, planes, first_slice;
  };
  bool shard_begin(uint32_t world, uint32_t rank) override {
    world_ = world;
    rank_ = rank;
    sharded_ = true;
    return true;
  }
  size_t shard_granularity() override { return 4096; }
  void* shard_alloc(size_t slice_bytes, size_t planes) override {
    void* va = mmap(nullptr, slice_bytes * planes * world_, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (va == MAP_FAILED) return nullptr;
    cols_.push_back(Col{(uint8_t*)va, slice_bytes, planes, n_slices_});
    n_slices_ += planes;
    return va;
  }
  bool map_slice(uint32_t r, size_t k, int fd) {
    for (const Col& c : cols_) {
      if (k < c.first_slice || k >= c.first_slice + c.planes) continue;
      uint8_t* at = c.va + ((k - c.first_slice) * world_ + r) * c.slice_bytes;
      return mmap(at, c.slice_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0) != MAP_FAILED;
    }
    return false;
  }
  bool shard_commit(const int** fds, size_t* n) override {
    fds_.assign(n_slices_, -1);
    for (const Col& c : cols_)
      for (size_t p = 0; p < c.planes; ++p) {
        int fd = memfd_create("gsim-hostemu-slice", 0);
        if (fd < 0 || ftruncate(fd, (off_t)c.slice_bytes) != 0) return false;
        fds_[c.first_slice + p] = fd;
        if (!map_slice(rank_, c.first_slice + p, fd)) return false;
      }
    *fds = fds_.data();
    *n = fds_.size();
    return true;
  }
  bool shard_attach(uint32_t peer, const int* fds, size_t n) override {
    if (n != n_slices_ || peer >= world_ || peer == rank_) return false;
    for (size_t k = 0; k < n; ++k)
      if (!map_slice(peer, k, fds[k])) return false;
    return true;
  }
  bool xbar_host(const GsXbar& xb) override {
    const uint32_t e = *xb.epoch + 1u;
    for (uint32_t r = 0; r < xb.world; ++r) __atomic_store_n(&xb.flags[r][xb.rank], e, __ATOMIC_RELEASE);
    for (uint32_t r = 0; r < xb.world; ++r)
      while ((int32_t)(__atomic_load_n(&xb.flags[xb.rank][r], __ATOMIC_ACQUIRE) - e) < 0) usleep(20);
    *xb.epoch = e;
    return true;
  }
  bool and_columns(const GsDev& d, const GsGlobals& g, uint32_t keep, uint32_t first, uint32_t count) override {
    for (uint32_t i = first; i < g.n && i - first < count; ++i) {
      d.heard[i] &= keep;
      d.queued[i] &= keep;
      for (uint32_t s = 0; s <= g.ring_mask; ++s) d.inbox[s][i] &= keep;
    }
    return true;
  }
  bool sync() override { return true; }
  const char* last_error() const override { return err_; }
  uint64_t total_launches() const override { return launches_; }

 private:
  int order_;
  bool no_fast_ = false;
  uint64_t launches_ = 0;
  bool sharded_ = false;
  uint32_t world_ = 1, rank_ = 0;
  size_t n_slices_ = 0;
  std::vector<Col> cols_;
  std::vector<int> fds_;
  char err_[128];
};

}  // namespace

GsBackend* gs_make_hostemu_backend(int, char*, size_t) { return new HostEmuBackend(); }
