// Register-ring streaming of per-level global operands through a k-recurrence.
#pragma once
#include "common.h"

namespace fv3 {

// k-loops over streamed global operands: body(k, a, b) for k in [kb, ke) ascending (or descending for stream_down),
// with a(k) = la(k), b(k) = lb(k) requested DEPTH trips before their use and held in a register ring.
#ifndef FV3_STREAM_DEPTH
#define FV3_STREAM_DEPTH 1  // measured on the remap chains (2.65 ms per call at depth 8): depth 1 2.28, 2 2.65, 3 2.60, 4 2.47, 6 2.57, 16 3.29 ms - the smallest code wins
#endif
#ifndef FV3_STREAM_DEPTH_DOWN
#define FV3_STREAM_DEPTH_DOWN FV3_STREAM_DEPTH
#endif
constexpr int STREAM_DEPTH = FV3_STREAM_DEPTH;
template <class LA, class LB, class B>
FV_DEV void stream_up(int kb, int ke, LA la, LB lb, B body) {
  double ra[STREAM_DEPTH], rb[STREAM_DEPTH];
#pragma unroll
  for (int n = 0; n < STREAM_DEPTH; ++n) {
    const int k = kb + n;
    ra[n] = k < ke ? la(k) : 0.0;
    rb[n] = k < ke ? lb(k) : 0.0;
  }
  for (int k0 = kb; k0 < ke; k0 += STREAM_DEPTH) {
#pragma unroll
    for (int n = 0; n < STREAM_DEPTH; ++n) {
      const int k = k0 + n;
      if (k < ke) {
        const double a = ra[n], b = rb[n];
        const int kn = k + STREAM_DEPTH;
        ra[n] = kn < ke ? la(kn) : 0.0;
        rb[n] = kn < ke ? lb(kn) : 0.0;
        body(k, a, b);
      }
    }
  }
}
constexpr int STREAM_DEPTH_DN = FV3_STREAM_DEPTH_DOWN;
template <class LA, class B>
FV_DEV void stream_down(int kb, int ke, LA la, B body) {  // k = ke-1 .. kb
  double ra[STREAM_DEPTH_DN];
#pragma unroll
  for (int n = 0; n < STREAM_DEPTH_DN; ++n) {
    const int k = ke - 1 - n;
    ra[n] = k >= kb ? la(k) : 0.0;
  }
  for (int k0 = ke - 1; k0 >= kb; k0 -= STREAM_DEPTH_DN) {
#pragma unroll
    for (int n = 0; n < STREAM_DEPTH_DN; ++n) {
      const int k = k0 - n;
      if (k >= kb) {
        const double a = ra[n];
        const int kn = k - STREAM_DEPTH_DN;
        ra[n] = kn >= kb ? la(kn) : 0.0;
        body(k, a);
      }
    }
  }
}


}  // namespace fv3
