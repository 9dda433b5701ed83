// Shared kernel scaffolding of the FV3 B200 hot path.
//
// Kernels are written as bulk-synchronous point/column functors ("FV_LAMBDA") launched through launch3d /
// launch2d.  Under nvcc they become __global__ kernels for sm_100a; compiled with -DFV3_HOSTSIM (g++, tests
// only) the same functors run in plain loops so host-side orchestration can be exercised on a CPU-only box.
// The product library never contains the host-simulation path (see pace_b200/_lib.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/fv3_b200.h"

#ifdef FV3_HOSTSIM
typedef void *cudaStream_t;
#define FV_HD inline
#define FV_DEV inline
#define FV_LDG(ptr) (*(ptr))
#define FV_LAMBDA [=]
#define FV_RESTRICT
#else
#include <cuda_runtime.h>
#define FV_HD __host__ __device__ __forceinline__
#define FV_DEV __device__ __forceinline__
#define FV_LDG(ptr) __ldg(ptr)  // read-only global operand: LDG.NC, free to be hoisted above shared-memory traffic
#define FV_LAMBDA [=] __device__
#define FV_RESTRICT __restrict__
#endif

struct fv3_ctx {
  fv3_geom g;
  fv3_config c;
  fv3_grid m;
  double *scratch;
  int64_t scratch_bytes;
  uint64_t uid;  // unique per fv3_create (a recycled address must not look like the previous context)
};

// Geometry and metric-term pointer tables live in __constant__ memory (one copy per translation unit, refreshed by
// the launchers whenever the active context changes): kernels read them as c[bank][offset] operands instead of
// carrying ~550 bytes of by-value closure that a dynamic index (edge[s]) would force into local memory.
#ifdef FV3_HOSTSIM
#define FV_DEV_GM
#else
static __constant__ fv3_geom c_g;
static __constant__ fv3_grid c_m;
#define FV_DEV_GM                 \
  const fv3_geom &g = c_g;        \
  const fv3_grid &m = c_m;        \
  (void)g;                        \
  (void)m;
#endif

namespace fv3 {

void set_error(const char *msg);
int check_launch(const char *what);
extern long long g_launches;  // kernels launched by this library since load (fv3_launch_count)

// scratch field n (3-D, all subdomains)
static inline double *scratch_field(const fv3_ctx *ctx, int n) {
  return ctx->scratch + (int64_t)n * ctx->g.ss * ctx->g.n_sub;
}

FV_HD bool on_west(const fv3_geom &g, int s) { return g.edge[s] & FV3_EDGE_WEST; }
FV_HD bool on_east(const fv3_geom &g, int s) { return g.edge[s] & FV3_EDGE_EAST; }
FV_HD bool on_south(const fv3_geom &g, int s) { return g.edge[s] & FV3_EDGE_SOUTH; }
FV_HD bool on_north(const fv3_geom &g, int s) { return g.edge[s] & FV3_EDGE_NORTH; }

FV_HD double dmin(double a, double b) { return a < b ? a : b; }
FV_HD double dmax(double a, double b) { return a > b ? a : b; }

#ifndef FV3_K3D_THREADS
#define FV3_K3D_THREADS 128
#endif
#ifndef FV3_HOSTSIM
template <class F>
__global__ void __launch_bounds__(FV3_K3D_THREADS) k3d(F f, int i0, int ni, int j0, int nj, int k0) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ni * nj) return;
  int j = idx / ni;
  int i = idx - j * ni;
  f((int)blockIdx.z, i0 + i, j0 + j, k0 + (int)blockIdx.y);
}
template <class F>
__global__ void __launch_bounds__(64) k2d(F f, int i0, int ni, int j0, int nj) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ni * nj) return;
  int j = idx / ni;
  int i = idx - j * ni;
  f((int)blockIdx.y, i0 + i, j0 + j);
}
template <class F>
__global__ void __launch_bounds__(128) k1d(F f, int64_t n) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) f(idx, (int)blockIdx.y, (int)blockIdx.z);
}
#endif

#ifndef FV3_HOSTSIM
static uint64_t tu_active_uid = 0;
static inline void activate(const fv3_ctx *ctx, cudaStream_t st) {
  if (tu_active_uid == ctx->uid) return;
  cudaMemcpyToSymbolAsync(c_g, &ctx->g, sizeof(fv3_geom), 0, cudaMemcpyHostToDevice, st);
  cudaMemcpyToSymbolAsync(c_m, &ctx->m, sizeof(fv3_grid), 0, cudaMemcpyHostToDevice, st);
  tu_active_uid = ctx->uid;
}
#endif

// f(s, i, j, k) for i in [i0, i1), j in [j0, j1), k in [k0, k1), all local subdomains
template <class F>
inline void launch3d(const fv3_ctx *ctx, cudaStream_t st, int i0, int i1, int j0, int j1, int k0, int k1, F f) {
  int ni = i1 - i0, nj = j1 - j0, nk = k1 - k0;
  if (ni <= 0 || nj <= 0 || nk <= 0) return;
#ifdef FV3_HOSTSIM
  (void)st;
  const int n_sub = ctx->g.n_sub;
#ifdef FV3_HOSTSIM_OMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
  for (int s = 0; s < n_sub; ++s)
    for (int k = k0; k < k1; ++k)
      for (int j = j0; j < j1; ++j)
        for (int i = i0; i < i1; ++i) f(s, i, j, k);
#else
  activate(ctx, st);
  dim3 grid((ni * nj + FV3_K3D_THREADS - 1) / FV3_K3D_THREADS, nk, ctx->g.n_sub);
  k3d<<<grid, FV3_K3D_THREADS, 0, st>>>(f, i0, ni, j0, nj, k0);
  ++g_launches;
#endif
}

// f(s, i, j): one thread per column
template <class F>
inline void launch2d(const fv3_ctx *ctx, cudaStream_t st, int i0, int i1, int j0, int j1, F f) {
  int ni = i1 - i0, nj = j1 - j0;
  if (ni <= 0 || nj <= 0) return;
#ifdef FV3_HOSTSIM
  (void)st;
  const int n_sub = ctx->g.n_sub;
#ifdef FV3_HOSTSIM_OMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
  for (int s = 0; s < n_sub; ++s)
    for (int j = j0; j < j1; ++j)
      for (int i = i0; i < i1; ++i) f(s, i, j);
#else
  activate(ctx, st);
  dim3 grid((ni * nj + 63) / 64, ctx->g.n_sub);
  k2d<<<grid, 64, 0, st>>>(f, i0, ni, j0, nj);
  ++g_launches;
#endif
}

// f(e, y, z) for e in [0, n), y in [0, ny), z in [0, nz)
template <class F>
inline void launch1d(cudaStream_t st, int64_t n, int ny, int nz, F f) {
  if (n <= 0 || ny <= 0 || nz <= 0) return;
#ifdef FV3_HOSTSIM
  (void)st;
#ifdef FV3_HOSTSIM_OMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int64_t e = 0; e < n; ++e) f(e, y, z);
#else
  dim3 grid((unsigned)((n + 127) / 128), ny, nz);
  k1d<<<grid, 128, 0, st>>>(f, n);
  ++g_launches;
#endif
}

}  // namespace fv3

// index helpers used inside functors (g = fv3_geom in scope)
#define O3(s, i, j, k) ((int64_t)(s) * g.ss + (int64_t)(k) * g.sk + (int64_t)(j) * g.sj + (i))
#define O2(s, i, j) ((int64_t)(s) * g.ss2 + (int64_t)(j) * g.sj + (i))
