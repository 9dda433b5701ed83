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

}  // extern "C"
