// gs_aux.h — per-row bodies of the auxiliary kernels (row init, crash injection,
// recount, state digest).  Same sharing arrangement as gs_row.h.
#pragma once
#include "gs_backend.h"
#include "gs_row.h"

// A converged member as serf.Create + a finished join leaves it ([U] memberlist.setAlive:
// incarnation 1; [U] serf.Create: the three Lamport clocks incremented once).
// Probe/gossip ticker phases mirror triggerFunc's random stagger ([U] state.go).
GS_DEV void gs_init_row(const GsDev& d, const GsGlobals& g, uint32_t i, uint32_t now) {
  const size_t cap = g.cap;
  // one stagger draw per phase group (default: per tile of 128 members, so that the
  // failure-detector path is uniform per CTA); phase_group == 1 draws it per member
  const uint32_t group = i / g.phase_group;
  const uint32_t pp = gs_probe_phase(g.rot_p, group, g.P), gp = gs_gossip_phase(g.rot_g, group, g.P, g.GI);
  const uint32_t k = gs_key_make(1u, 0u, GS_RANK_ALIVE, GS_TRUTH_UP);
  gs_key_store(d, g, 0u, i, k);
  gs_key_store(d, g, 1u, i, k);
  for (uint32_t s = 0; s <= g.ring_mask; ++s) d.inbox[s][i] = 0u;
  d.meta[i] = gp << GS_META_GPHASE_SHIFT;
  d.due[i] = now + (pp + g.P - now % g.P) % g.P;  // first tick >= now congruent to the phase
  d.cursor[i] = 0u;
  d.pass[i] = 0u;
  d.probe_tgt[i] = 0u;
  d.probe_inc[i] = 0u;
  d.sus_start[i] = 0u;
  for (uint32_t q = 0; q < GS_K1MAX; ++q) {
    d.sus_from[(size_t)q * cap + i] = GS_EMPTY32;
    d.acc[(size_t)q * cap + i] = GS_EMPTY64;
    d.acc[((size_t)GS_K1MAX + q) * cap + i] = GS_EMPTY64;
  }
  d.change_tick[i] = 0u;
  d.reap_after[i] = 0u;
  d.ltime_member[i] = 1u;
  d.ltime_event[i] = 1u;
  d.event_min[i] = 0u;
  d.heard[i] = 0u;
  d.queued[i] = 0u;
  if (d.coord != nullptr) {  // [U] coordinate.NewCoordinate: the origin, maximal error, minimal height
    GsCoord o;
    gs_coord_origin(o);
    for (uint32_t slot = 0; slot < 2u; ++slot) {
      double* out = d.coord + ((size_t)slot * GS_COORD_WORDS) * cap + i;
      for (uint32_t x = 0; x < GS_COORD_DIM; ++x) out[(size_t)x * cap] = o.vec[x];
      out[(size_t)8 * cap] = o.error;
      out[(size_t)9 * cap] = o.adjustment;
      out[(size_t)10 * cap] = o.height;
      d.ctag[(size_t)slot * cap + i] = 0u;
    }
    for (uint32_t s = 0; s < GS_ADJ_WINDOW; ++s) d.adj[(size_t)s * cap + i] = 0.0;
    d.adj_idx[i] = 0u;
  }
  if (d.ppreq != nullptr) {
    for (uint32_t q = 0; q < 2u * GS_PPK; ++q) d.ppreq[(size_t)q * cap + i] = GS_EMPTY32;
    for (uint32_t q = 0; q < 4u; ++q) d.pp_clk[(size_t)q * cap + i] = 0u;
  }
}

// BASELINE config 3: crash every UP member whose Philox draw is below the threshold.
GS_DEV bool gs_crash_row(const GsDev& d, const GsGlobals& g, uint32_t i, uint32_t thr,
                         uint32_t salt) {
  uint32_t k = d.key[0][i];
  if (gs_key_truth(k) != GS_TRUTH_UP) return false;
  GsU4 r = gs_philox(g.seed_lo, g.seed_hi, i, salt, GS_PUR_CRASH, 0u);
  if (r.x >= thr) return false;
  k = (k & ~3u) | GS_TRUTH_CRASHED;
  gs_key_store(d, g, 0u, i, k);
  gs_key_store(d, g, 1u, i, (d.key[1][i] & ~3u) | GS_TRUTH_CRASHED);
  return true;
}

// [U] serf.handleReap -> reap(failedMembers, ReconnectTimeout) / reap(leftMembers, TombstoneTimeout):
// a member that has been Failed (Left) for longer than the timeout is erased from the member
// list (EventMemberReap).  Row a17 of SURVEY 8a; Consul shortens the timeouts in
// agent/consul/server_test.go:675-677.  Returns bit 0 = reaped, bit 1 = it was an established
// (non-pending) member; the caller logs the event.
GS_DEV uint32_t gs_reap_row(const GsDev& d, const GsGlobals& g, uint32_t i, uint32_t now,
                            uint32_t reconnect_ticks, uint32_t tombstone_ticks) {
  const uint32_t k = d.key[now & 1u][i];
  const uint32_t truth = gs_key_truth(k), rank = gs_key_rank(k);
  if (truth == GS_TRUTH_NONE || truth == GS_TRUTH_UP) return 0u;
  uint32_t limit;
  if (rank == GS_RANK_DEAD) limit = d.reap_after[i] != 0u ? d.reap_after[i] : reconnect_ticks;  // ReconnectTimeoutOverride
  else if (rank == GS_RANK_LEFT) limit = tombstone_ticks;
  else return 0u;
  if (now - d.change_tick[i] <= limit) return 0u;
  gs_key_store(d, g, 0u, i, d.key[0][i] & ~3u);
  gs_key_store(d, g, 1u, i, d.key[1][i] & ~3u);
  return 1u | (gs_key_pending(k) ? 0u : 2u);
}

// Canonical digest of one row: only fields that are semantically live are folded, so
// that stale scratch in cold columns never matters (the oracle folds the same fields).
GS_DEV uint64_t gs_hash_row(const GsDev& d, const GsGlobals& g, uint32_t i, uint32_t now) {
  const uint32_t cur = now & 1u;  // buffer the next tick will read
  const size_t cap = g.cap;
  const uint32_t k = d.key[cur][i];
  const uint32_t truth = gs_key_truth(k);
  if (truth == GS_TRUTH_NONE) return 0ull;
  const uint32_t m = d.meta[i] & ~GS_META_DIRTY;
  const uint32_t rank = gs_key_rank(k);
  const bool up = truth == GS_TRUTH_UP;
  const bool probing = up && gs_meta_stage(m) != GS_STAGE_IDLE;
  uint64_t h = 0x9E3779B97F4A7C15ull;
  h = gs_mix64(h, i);
  h = gs_mix64(h, k);
  h = gs_mix64(h, m);
  h = gs_mix64(h, up ? d.due[i] : 0u);
  h = gs_mix64(h, d.cursor[i]);
  h = gs_mix64(h, d.pass[i]);
  h = gs_mix64(h, probing ? d.probe_tgt[i] : 0u);
  h = gs_mix64(h, probing ? d.probe_inc[i] : 0u);
  h = gs_mix64(h, rank == GS_RANK_SUSPECT ? d.sus_start[i] : 0u);
  for (uint32_t q = 0; q < GS_K1MAX; ++q)
    h = gs_mix64(h, rank == GS_RANK_SUSPECT ? d.sus_from[(size_t)q * cap + i] : 0u);
  h = gs_mix64(h, rank >= GS_RANK_DEAD ? d.change_tick[i] : 0u);
  h = gs_mix64(h, d.ltime_member[i]);
  h = gs_mix64(h, d.ltime_event[i]);
  h = gs_mix64(h, d.event_min[i]);
  const uint32_t heard = d.heard[i] & g.active_mask;
  h = gs_mix64(h, heard);
  h = gs_mix64(h, d.queued[i] & g.active_mask);
  const uint32_t inb = d.inbox[now & g.ring_mask][i];
  h = gs_mix64(h, inb & (g.active_mask | GS_ACC_BIT));
  // latency pools: packets still in flight, by ticks until arrival (slot now-1 was just consumed)
  for (uint32_t s = 1; s < g.ring_mask; ++s)
    h = gs_mix64(h, d.inbox[(now + s) & g.ring_mask][i] & g.active_mask);
  uint32_t hm = heard;
  while (hm) {
#if defined(__CUDA_ARCH__)
    uint32_t r = __ffs(hm) - 1;
#else
    uint32_t r = (uint32_t)__builtin_ctz(hm);
#endif
    hm &= hm - 1;
    h = gs_mix64(h, (r << 8) | d.tx[GS_TX(r, cap, i)]);
  }
  if (d.coord != nullptr) {  // the member's current coordinate, bit for bit, and its adjustment window
    const uint32_t slot = d.ctag[cap + i] > d.ctag[i] ? 1u : 0u;
    const double* c = d.coord + ((size_t)slot * GS_COORD_WORDS) * cap + i;
    for (uint32_t x = 0; x < GS_COORD_WORDS; ++x) {
      uint64_t bits;
      memcpy(&bits, &c[(size_t)x * cap], 8);
      h = gs_mix64(h, bits);
    }
    h = gs_mix64(h, d.adj_idx[i]);
  }
  if (inb & GS_ACC_BIT) {
    const uint64_t* acc = d.acc + (size_t)cur * GS_K1MAX * cap;
    for (uint32_t s = 0; s < GS_K1MAX; ++s) h = gs_mix64(h, acc[(size_t)s * cap + i]);
    if (g.pp_interval != 0u) {  // push-pull requests and partner clocks waiting in the mailbox
      const uint32_t* req = d.ppreq + (size_t)cur * GS_PPK * cap;
      for (uint32_t s = 0; s < GS_PPK; ++s) h = gs_mix64(h, req[(size_t)s * cap + i]);
      const uint32_t* clk = d.pp_clk + (size_t)cur * 2u * cap;
      h = gs_mix64(h, clk[i]);
      h = gs_mix64(h, clk[cap + i]);
    }
  }
  return h;
}

