// Register-ring streaming of per-level global operands through a k-recurrence.
#pragma once
#include "common.h"

namespace fv3 {

// k-loops over streamed global operands: body(k, a, b) for k in [kb, ke) ascending (or descending for stream_down),
// with a(k) = la(k), b(k) = lb(k) requested DEPTH trips before their use and held in a register ring.
constexpr int STREAM_DEPTH = 8;
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
template <class LA, class B>
FV_DEV void stream_down(int kb, int ke, LA la, B body) {  // k = ke-1 .. kb
  double ra[STREAM_DEPTH];
#pragma unroll
  for (int n = 0; n < STREAM_DEPTH; ++n) {
    const int k = ke - 1 - n;
    ra[n] = k >= kb ? la(k) : 0.0;
  }
  for (int k0 = ke - 1; k0 >= kb; k0 -= STREAM_DEPTH) {
#pragma unroll
    for (int n = 0; n < STREAM_DEPTH; ++n) {
      const int k = k0 - n;
      if (k >= kb) {
        const double a = ra[n];
        const int kn = k - STREAM_DEPTH;
        ra[n] = kn >= kb ? la(kn) : 0.0;
        body(k, a);
      }
    }
  }
}


}  // namespace fv3
