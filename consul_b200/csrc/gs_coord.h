// gs_coord.h — network coordinates (Vivaldi), SURVEY 8f N3 / 8b GetCoordinate.
//
// [U] serf/coordinate/{config,coordinate,client}.go: every direct ack of a probe carries the
// target's coordinate; the prober moves its own coordinate along the error-weighted spring force
// between the measured and the predicted round trip, keeps a 20-sample adjustment term, and is
// pulled weakly towards the origin.  Consul reads the result through (*Serf).GetCoordinate /
// GetCachedCoordinate (agent/router/router.go:62-67) to sort datacenters and servers by distance.
//
// Opt-in (GSIM_FLAG_COORDINATES).  What is NOT restated: the per-peer latency filter (a median
// of the last 3 samples per (observer, peer) pair — O(N^2) state); samples are used as measured.
// Arithmetic is IEEE double in the order written here; the oracle repeats it independently and
// the build disables FMA contraction on both sides, so results are compared bit for bit.
//
// Publication without copying: each member owns two slots tagged with (tick written + 1).  The
// owner always overwrites the OLDER slot; a reader at tick t takes the newer slot among those
// with tag <= t.  The slot being written in tick t therefore is never the one readers use.
#pragma once
#include <math.h>

#include "gs_core.h"

#define GS_COORD_DIM 8
#define GS_COORD_WORDS 11  // vec[8], error, adjustment, height
#define GS_ADJ_WINDOW 20
#define GS_PUR_COORD 8

struct GsCoord {
  double vec[GS_COORD_DIM];
  double error, adjustment, height;
};

#define GS_VIVALDI_ERROR_MAX 1.5
#define GS_VIVALDI_CE 0.25
#define GS_VIVALDI_CC 0.25
#define GS_HEIGHT_MIN 10.0e-6
#define GS_GRAVITY_RHO 150.0
#define GS_ZERO_THRESHOLD 1.0e-6

GS_HD void gs_coord_origin(GsCoord& c) {
  for (int x = 0; x < GS_COORD_DIM; ++x) c.vec[x] = 0.0;
  c.error = GS_VIVALDI_ERROR_MAX;
  c.adjustment = 0.0;
  c.height = GS_HEIGHT_MIN;
}

GS_HD double gs_coord_magnitude(const double* v) {
  double sum = 0.0;
  for (int x = 0; x < GS_COORD_DIM; ++x) sum += v[x] * v[x];
  return sqrt(sum);
}

GS_HD double gs_coord_raw_distance(const GsCoord& a, const GsCoord& b) {
  double d[GS_COORD_DIM];
  for (int x = 0; x < GS_COORD_DIM; ++x) d[x] = a.vec[x] - b.vec[x];
  return gs_coord_magnitude(d) + a.height + b.height;
}

// Coordinate.DistanceTo(other).Seconds(): the adjusted distance, through time.Duration (whole
// nanoseconds, truncated) and back.
GS_HD double gs_coord_distance_seconds(const GsCoord& a, const GsCoord& b) {
  double dist = gs_coord_raw_distance(a, b);
  const double adjusted = dist + a.adjustment + b.adjustment;
  if (adjusted > 0.0) dist = adjusted;
  const long long ns = (long long)(dist * 1.0e9);
  return (double)(ns / 1000000000LL) + (double)(ns % 1000000000LL) / 1.0e9;
}

// Coordinate.ApplyForce: move `c` by `force` along the unit vector from `other` to `c`; two
// coincident points are pushed apart in a pseudo-random direction (Philox instead of math/rand).
GS_HD void gs_coord_apply_force(GsCoord& c, double force, const GsCoord& other, uint32_t seed_lo,
                                uint32_t seed_hi, uint32_t member, uint32_t tick, uint32_t salt) {
  double unit[GS_COORD_DIM];
  for (int x = 0; x < GS_COORD_DIM; ++x) unit[x] = c.vec[x] - other.vec[x];
  double mag = gs_coord_magnitude(unit);
  if (mag > GS_ZERO_THRESHOLD) {
    const double inv = 1.0 / mag;
    for (int x = 0; x < GS_COORD_DIM; ++x) unit[x] = unit[x] * inv;
  } else {
    const GsU4 r0 = gs_philox(seed_lo, seed_hi, member, tick, GS_PUR_COORD, salt * 2u);
    const GsU4 r1 = gs_philox(seed_lo, seed_hi, member, tick, GS_PUR_COORD, salt * 2u + 1u);
    const uint32_t w[GS_COORD_DIM] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    for (int x = 0; x < GS_COORD_DIM; ++x) unit[x] = (double)w[x] / 4294967296.0 - 0.5;
    const double m2 = gs_coord_magnitude(unit);
    if (m2 > GS_ZERO_THRESHOLD) {
      const double inv = 1.0 / m2;
      for (int x = 0; x < GS_COORD_DIM; ++x) unit[x] = unit[x] * inv;
    } else {
      for (int x = 0; x < GS_COORD_DIM; ++x) unit[x] = 0.0;
      unit[0] = 1.0;
    }
    mag = 0.0;
  }
  for (int x = 0; x < GS_COORD_DIM; ++x) c.vec[x] = c.vec[x] + unit[x] * force;
  if (mag > GS_ZERO_THRESHOLD) {
    c.height = (c.height + other.height) * force / mag + c.height;
    if (!(c.height > GS_HEIGHT_MIN)) c.height = GS_HEIGHT_MIN;  // math.Max(height, HeightMin)
  }
}

GS_HD bool gs_coord_valid(const GsCoord& c) {
  bool ok = isfinite(c.error) && isfinite(c.adjustment) && isfinite(c.height);
  for (int x = 0; x < GS_COORD_DIM; ++x) ok = ok && isfinite(c.vec[x]);
  return ok;
}

// Client.Update(node, other, rtt) minus the latency filter.  `samples` = this member's adjustment
// window (GS_ADJ_WINDOW doubles, strided), `idx` its write position.
GS_HD void gs_coord_client_update(GsCoord& c, const GsCoord& other, double rtt_seconds, double* samples,
                                  size_t sample_stride, uint32_t* idx, uint32_t seed_lo, uint32_t seed_hi,
                                  uint32_t member, uint32_t tick) {
  // updateVivaldi
  const double dist = gs_coord_distance_seconds(c, other);
  if (rtt_seconds < GS_ZERO_THRESHOLD) rtt_seconds = GS_ZERO_THRESHOLD;
  const double wrongness = fabs(dist - rtt_seconds) / rtt_seconds;
  double total_error = c.error + other.error;
  if (total_error < GS_ZERO_THRESHOLD) total_error = GS_ZERO_THRESHOLD;
  const double weight = c.error / total_error;
  c.error = GS_VIVALDI_CE * weight * wrongness + c.error * (1.0 - GS_VIVALDI_CE * weight);
  if (c.error > GS_VIVALDI_ERROR_MAX) c.error = GS_VIVALDI_ERROR_MAX;
  const double delta = GS_VIVALDI_CC * weight;
  const double force = delta * (rtt_seconds - dist);
  gs_coord_apply_force(c, force, other, seed_lo, seed_hi, member, tick, 0u);
  // updateAdjustment
  const double raw = gs_coord_raw_distance(c, other);
  samples[(size_t)(*idx) * sample_stride] = rtt_seconds - raw;
  *idx = (*idx + 1u) % GS_ADJ_WINDOW;
  double sum = 0.0;
  for (uint32_t s = 0; s < GS_ADJ_WINDOW; ++s) sum += samples[(size_t)s * sample_stride];
  c.adjustment = sum / (2.0 * (double)GS_ADJ_WINDOW);
  // updateGravity
  GsCoord origin;
  gs_coord_origin(origin);
  const double od = gs_coord_distance_seconds(origin, c);
  const double q = od / GS_GRAVITY_RHO;
  const double gforce = -1.0 * (q * q);
  gs_coord_apply_force(c, gforce, origin, seed_lo, seed_hi, member, tick, 1u);
  if (!gs_coord_valid(c)) gs_coord_origin(c);  // Client.Update resets an invalid coordinate
}
