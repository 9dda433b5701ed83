// Range / NaN check of a state field: minimum, maximum and NaN count in ONE pass over the field, for the optional sanity
// checks a driver runs every few steps.
//   fv3_field_check <- SafetyChecker.check_state (driver/pace/driver/safety_checks.py:70-110: min / max over the compute
//                      or the whole domain, NaN test) and the negative-delp / negative-tracer / NaN checks DaCe debug
//                      passes inject (dsl/pace/dsl/dace/sdfg_debug_passes.py:185-269)
#include <cstdint>

#include "common.h"

namespace {
// order-preserving map of a (non-NaN) double to a signed 64-bit key, so that min / max are integer atomics
FV_HD long long key_of(double v) {
  long long b;
  memcpy(&b, &v, 8);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
}  // namespace

extern "C" {

// out: three 64-bit words on the device — key of the minimum, key of the maximum (fv3 key: bits ^ ((bits >> 63) &
// 0x7fff...f), its own inverse) over the non-NaN values, and the number of NaN values.  Points i in [i0, i1), j in
// [j0, j1) of levels [0, nk) of every local subdomain; nk == 0: the field is 2-D ([s][j][i]).
int fv3_field_check(fv3_ctx *ctx, const double *field, int i0, int i1, int j0, int j1, int nk, int64_t *out, void *stream) {
  const fv3_geom g = ctx->g;
  if (nk < 0 || nk > g.nk || i0 < 0 || i1 > g.ni || j0 < 0 || j1 > g.nj || i1 <= i0 || j1 <= j0) {
    fv3::set_error("fv3_field_check: range outside the storage");
    return -1;
  }
  const bool two_d = nk == 0;
  if (two_d) nk = 1;
  long long *o = (long long *)out;
#ifdef FV3_HOSTSIM
  (void)stream;
  long long kmin = INT64_MAX, kmax = INT64_MIN, nan = 0;
  for (int s = 0; s < g.n_sub; ++s)
    for (int k = 0; k < nk; ++k)
      for (int j = j0; j < j1; ++j)
        for (int i = i0; i < i1; ++i) {
          const double v = field[two_d ? O2(s, i, j) : O3(s, i, j, k)];
          if (v != v) {
            ++nan;
            continue;
          }
          const long long key = key_of(v);
          kmin = key < kmin ? key : kmin;
          kmax = key > kmax ? key : kmax;
        }
  o[0] = kmin;
  o[1] = kmax;
  o[2] = nan;
  return 0;
#else
  cudaStream_t st = (cudaStream_t)stream;
  fv3::launch1d(st, 1, 1, 1, FV_LAMBDA(int64_t, int, int) {
    o[0] = INT64_MAX;
    o[1] = INT64_MIN;
    o[2] = 0;
  });
  // one thread per column position and level; a warp combines its values by shuffles and issues at most three atomics
  fv3::launch3d(ctx, st, i0, i1, j0, j1, 0, nk, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const double v = field[two_d ? O2(s, i, j) : O3(s, i, j, k)];
    const bool isn = v != v;
    long long kmin = isn ? INT64_MAX : key_of(v), kmax = isn ? INT64_MIN : key_of(v);
    const unsigned act = __activemask();
    unsigned nn = __popc(__ballot_sync(act, isn));
    for (int d = 16; d > 0; d >>= 1) {
      const long long a = __shfl_xor_sync(act, kmin, d), b = __shfl_xor_sync(act, kmax, d);
      // a lane outside the active mask returns an undefined value: only combine with active partners
      const bool ok = (act >> ((threadIdx.x & 31) ^ d)) & 1u;
      if (ok) {
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
      }
    }
    // after the xor butterfly every lane of a FULL warp holds the warp result; a partial warp (the ragged end of a
    // row block) lets every lane contribute its partial value — still correct, just more atomics
    const bool full = act == 0xffffffffu;
    if (!full || (threadIdx.x & 31) == 0) {
      if (kmin != INT64_MAX) atomicMin(o, kmin);
      if (kmax != INT64_MIN) atomicMax(o + 1, kmax);
    }
    if ((threadIdx.x & 31) == (unsigned)(__ffs(act) - 1) && nn) atomicAdd((unsigned long long *)(o + 2), (unsigned long long)nn);
  });
  return fv3::check_launch("fv3_field_check");
#endif
}

}  // extern "C"
