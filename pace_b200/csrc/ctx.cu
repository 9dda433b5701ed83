// Context, error reporting and ABI bookkeeping of libfv3b200.
#include "common.h"

namespace fv3 {
static char g_err[512] = "";
long long g_launches = 0;
void set_error(const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); }
int check_launch(const char *what) {
#ifdef FV3_HOSTSIM
  (void)what;
  return 0;
#else
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
#endif
}
}  // namespace fv3

extern "C" {

const char *fv3_last_error(void) { return fv3::g_err; }
int fv3_abi_version(void) { return 1; }
int fv3_is_hostsim(void) {
#ifdef FV3_HOSTSIM
  return 1;
#else
  return 0;
#endif
}
int fv3_scratch_fields(void) { return 40; }
int64_t fv3_launch_count(void) { return (int64_t)fv3::g_launches; }

fv3_ctx *fv3_create(const fv3_geom *geom, const fv3_config *config, const fv3_grid *grid, void *scratch,
                    int64_t scratch_bytes) {
  if (!geom || !config || !grid) {
    fv3::set_error("fv3_create: null argument");
    return nullptr;
  }
  if (geom->n_sub < 1 || geom->n_sub > FV3_MAX_SUBDOMAINS || geom->nk > FV3_MAX_LEVELS) {
    fv3::set_error("fv3_create: n_sub or nk out of range");
    return nullptr;
  }
  int64_t need = (int64_t)fv3_scratch_fields() * geom->ss * geom->n_sub * 8;
  if (scratch_bytes < need || !scratch) {
    fv3::set_error("fv3_create: scratch buffer too small");
    return nullptr;
  }
  fv3_ctx *ctx = new fv3_ctx;
  ctx->g = *geom;
  ctx->c = *config;
  ctx->m = *grid;
  ctx->scratch = (double *)scratch;
  ctx->scratch_bytes = scratch_bytes;
  static uint64_t next_uid = 0;
  ctx->uid = ++next_uid;
  return ctx;
}

void fv3_destroy(fv3_ctx *ctx) { delete ctx; }

// Measurement aid (bench.py): a dependent-chain-free stream of fp64 FMAs on every SM, `chains` independent accumulators
// per thread.  2 * iters * chains * threads flops per launch; out receives one value per thread so that nothing is
// optimised away.  Gives the fp64 roof the plane / column kernels are compared with (they are instruction-bound, not
// HBM-bound).
int fv3_fp64_peak(double *out, int iters, int blocks, void *stream) {
#ifdef FV3_HOSTSIM
  (void)out; (void)iters; (void)blocks; (void)stream;
  return 0;
#else
  fv3::launch1d((cudaStream_t)stream, (int64_t)blocks * 128, 1, 1, FV_LAMBDA(int64_t e, int, int) {
    double a0 = 1.0 + 1e-9 * (double)e, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0,
           a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 1.0 - 1e-12, c = 1e-13;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
      a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    out[e] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  });
  return fv3::check_launch("fv3_fp64_peak");
#endif
}

}  // extern "C"
