// gs_backend.h — the narrow device interface the host side of libgsim drives.
// libgsim.so links exactly one implementation: the CUDA backend (gs_cuda.cu).  A second
// implementation exists only under tests/hostemu/ (the same row function compiled by g++
// and looped sequentially) so kernel logic can be debugged in a GPU-less container; it is
// test infrastructure and is never linked into or loaded by the product library.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "gs_core.h"

struct GsRecount {
  uint32_t heard_cnt[GS_MAX_RUMORS];
  uint32_t queued_cnt[GS_MAX_RUMORS];
  uint32_t truth_cnt[4];
  uint32_t rank_cnt[4];
  uint32_t crashed_alive;
  uint32_t isolated_up;  // running members that have not joined the established set
  uint32_t pending;           // members not yet folded into the established set (known through their alive rumor only)
  uint32_t unreachable_live;  // members whose process is gone (crashed / shut down) but whom the cluster
                              // still lists as alive or suspect: the targets an unanswered probe can hit
};

class GsBackend {
 public:
  virtual ~GsBackend() {}
  virtual const char* name() const = 0;
  virtual void* alloc(size_t bytes) = 0;  // returns nullptr on failure
  virtual void release(void* p) = 0;
  virtual bool h2d(void* dst, const void* src, size_t bytes) = 0;
  virtual bool d2h(void* dst, const void* src, size_t bytes) = 0;
  // A few bytes host -> device, ordered on the pool's stream but NOT waited for: the source may be reused
  // as soon as the call returns (the driver stages small pageable copies before returning), everything
  // the pool launches or reads later comes after it.  The host side of Join / UserEvent is dozens of
  // single-word writes; waiting for each one was most of its cost.
  virtual bool h2d_word(void* dst, const void* src, size_t bytes) { return h2d(dst, src, bytes); }
  // Bulk host -> device, enqueued only: the caller keeps `src` alive and calls sync() before it returns
  // (gsim_restore streams its planes back to back and waits once).
  // host staging memory for bulk copies (page-locked where that makes the copy a plain DMA)
  virtual void* host_alloc(size_t bytes) { return malloc(bytes); }
  virtual void host_free(void* q) { free(q); }
  virtual bool h2d_async(void* dst, const void* src, size_t bytes) { return h2d(dst, src, bytes); }
  // The per-member words the host side of a state exchange needs, in one round trip:
  // out = {key[0], key[1], meta, heard, queued, ltime_member, ltime_event, event_min}
  virtual bool row_read(const GsDev& d, uint32_t i, uint32_t out[8]) = 0;
  virtual bool fill32(uint32_t* dst, uint32_t value, size_t count) = 0;
  virtual bool fill8(uint8_t* dst, uint8_t value, size_t count) = 0;
  // rows [first, first+count): converged members, inc=1, clocks=1, phases from Philox
  virtual bool init_rows(const GsDev& d, const GsGlobals* g_dev, const GsGlobals& g, uint32_t first,
                         uint32_t count, uint32_t now) = 0;
  // advance `nticks` ticks starting at tick t0 (tick_base on the device == t0 on entry and
  // t0+nticks on exit).  kernel_ms accumulates CUDA-event time of the tick launches.
  virtual bool run_ticks(const GsDev& d, const GsGlobals* g_dev, const GsGlobals& g, uint32_t t0,
                         uint32_t nticks, bool use_graph, double* kernel_ms, uint64_t* launches,
                         const GsXbar* xbar = nullptr) = 0;
  // Quiet windows: advance up to `nticks` ticks starting at t0 as a chain of launches
  // of <= ProbeInterval ticks each, on a pool whose mailboxes are known to be empty.  The chain stops at
  // the horizon (GS_Q_HORIZON); *ticks_done = how far it got (tick_base == t0 + *ticks_done on exit).
  // `per_launch` = ticks one launch may cover: ProbeInterval in general; more when the caller knows that
  // no probe can go unanswered (every listed member runs, no loss, no slow link), so the horizon cannot move.
  virtual bool run_windows(const GsDev& d, const GsGlobals* g_dev, const GsGlobals& g, uint32_t t0, uint32_t nticks,
                           uint32_t per_launch, bool use_graph, double* kernel_ms, uint64_t* launches,
                           uint32_t* ticks_done, const GsXbar* xbar, bool pristine) = 0;
  // lowers GS_Q_HORIZON (every rank's copy) to the earliest accusation the probes in flight of rows
  // [first, first+count) can produce
  virtual bool quiet_scan(const GsDev& d, const GsGlobals* g_dev, const GsGlobals& g, uint32_t now, uint32_t first,
                          uint32_t count) = 0;
  // ---- sharded (multi-GPU) pools, see gs_vmm.h; unsupported by default ------------------------
  virtual bool shard_begin(uint32_t, uint32_t) { return false; }
  virtual size_t shard_granularity() { return 0; }
  virtual void* shard_alloc(size_t /*slice_bytes*/, size_t /*planes*/) { return nullptr; }
  virtual bool shard_commit(const int** /*fds*/, size_t* /*n*/) { return false; }
  virtual bool shard_attach(uint32_t /*peer*/, const int* /*fds*/, size_t /*n*/) { return false; }
  virtual bool xbar_host(const GsXbar&) { return false; }  // arrive, wait for every rank, sync
  virtual bool crash_fraction(const GsDev& d, const GsGlobals* g_dev, const GsGlobals& g,
                              uint32_t thr, uint32_t salt, uint32_t now, uint32_t* n_crashed) = 0;
  // counts over members [first, first + count) (a rank of a sharded pool counts its own rows)
  virtual bool recount(const GsDev& d, const GsGlobals* g_dev, const GsGlobals& g, uint32_t now, uint32_t first,
                       uint32_t count, GsRecount* out) = 0;
  virtual bool state_hash(const GsDev& d, const GsGlobals* g_dev, const GsGlobals& g, uint32_t now,
                          uint64_t out[4]) = 0;
  // serf's reaper (gs_aux.h gs_reap_row) over every row; counts[0] = members erased, counts[1] =
  // how many of them were established; logs EventMemberReap when `log_events`
  virtual bool reap_rows(const GsDev& d, const GsGlobals* g_dev, const GsGlobals& g, uint32_t now,
                         uint32_t reconnect_ticks, uint32_t tombstone_ticks, bool log_events,
                         uint32_t counts[2]) = 0;
  // clear rumor bits outside `keep` in the heard / queued / mailbox columns (slot retirement)
  virtual bool and_columns(const GsDev& d, const GsGlobals& g, uint32_t keep, uint32_t first, uint32_t count) = 0;
  virtual bool sync() = 0;
  virtual const char* last_error() const = 0;
  virtual uint64_t total_launches() const = 0;
};

// Implemented by gs_cuda.cu (product) — returns nullptr and fills err when no usable
// sm_100-class device exists.  There is deliberately no CPU implementation in libgsim.
GsBackend* gs_make_cuda_backend(int device, char* err, size_t err_cap);

// ---- state digest shared by every implementation of state_hash -----------------
GS_HD uint64_t gs_mix64(uint64_t h, uint64_t w) {
  h = (h ^ w) * 0xff51afd7ed558ccdull;
  h ^= h >> 32;
  return h;
}
GS_HD void gs_hash_lanes(uint64_t h, uint64_t lanes[4]) {
  lanes[0] = h;
  lanes[1] = gs_mix64(h, 1);
  lanes[2] = gs_mix64(h, 2);
  lanes[3] = gs_mix64(h, 3);
}
