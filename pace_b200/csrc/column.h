// Column-tile kernels: the vertical implicit solvers and the vertical remap.
//
// One CTA owns a tile of COL_TILE = 32 horizontally adjacent columns (all levels) of one subdomain and keeps the
// per-level intermediates of those columns in shared memory as [level][column] arrays (conflict-free: lane = column).
// A stage is a sequence of block-wide phases of two kinds:
//   levels(k0, k1, f)   f(k, c): level-parallel work (logs, exps, divides; no dependence along k) spread over all
//                       warps of the CTA — warp w takes levels k0+w, k0+w+W, ...; lane = column
//   columns(f)          f(c): the k-recurrences (cumulative sums, Thomas forward/backward sweeps), one thread per
//                       column (warp 0), operands read from / written to the shared arrays
// so nothing spills to thread-local memory, every global field is read / written exactly once with coalesced
// accesses, and several CTAs per SM overlap one tile's recurrence with another tile's level-parallel math.
//
//   launch_columns(ctx, st, i0, i1, j0, j1, n_arrays, f)    f(Tile&) runs once per tile
//
// Under -DFV3_HOSTSIM (g++, tests only) a tile is a plain loop nest, so the same phase code runs on the CPU.
#pragma once
#include <vector>

#include "common.h"

namespace fv3 {

constexpr int COL_TILE = 32;
// Warps per tile and tiles per SM (measured on the Riemann solvers, whose three [level][column] arrays take 61 KB per tile:
// 2 x 16 warps 850 us, 3 x 14 760 us, 3 x 12 705 us, 3 x 10 775 us, 3 x 8 810 us per call at C128)
#ifndef FV3_COL_WARPS
#define FV3_COL_WARPS 12
#endif
#ifndef FV3_COL_MINB
#define FV3_COL_MINB 3
#endif
constexpr int COL_WARPS = FV3_COL_WARPS;

struct Tile {
  double *sm;       // n_arrays * nlev * COL_TILE doubles
  int s;            // subdomain
  int i0, ni, j0;   // column domain of the launch: i in [i0, i0+ni), rows from j0
  int first, ncol;  // first flattened column index of this tile, number of valid columns in it (<= COL_TILE)
  int nlev;         // levels per array (nz + 1)

  // array a, level k, column c
  FV_HD double *arr(int a) const { return sm + (size_t)a * nlev * COL_TILE; }
  // horizontal position of column c of this tile
  FV_HD void ij(int c, int &i, int &j) const {
    const int idx = first + c;
    const int jr = idx / ni;
    i = i0 + (idx - jr * ni);
    j = j0 + jr;
  }
#ifdef FV3_HOSTSIM
  template <class F>
  void levels(int k0, int k1, F f) const {
    for (int k = k0; k < k1; ++k)
      for (int c = 0; c < ncol; ++c) f(k, c);
  }
  template <class F>
  void columns(F f) const {
    for (int c = 0; c < ncol; ++c) f(c);
  }
#else
  template <class F>
  __device__ __forceinline__ void levels(int k0, int k1, F f) const {
    const int c = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (c < ncol)
      for (int k = k0 + w; k < k1; k += 2 * nw) {  // two levels per trip: their global loads issue back to back
        f(k, c);
        if (k + nw < k1) f(k + nw, c);
      }
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void columns(F f) const {
    if (threadIdx.x < ncol) f((int)threadIdx.x);
    __syncthreads();
  }
#endif
};

#ifndef FV3_HOSTSIM
template <class F>
__global__ void __launch_bounds__(COL_TILE *COL_WARPS, FV3_COL_MINB) kcolumns(F f, int i0, int ni, int j0, int ncols, int nlev) {
  extern __shared__ double col_smem[];
  Tile t;
  t.sm = col_smem;
  t.s = (int)blockIdx.y;
  t.i0 = i0;
  t.ni = ni;
  t.j0 = j0;
  t.first = (int)blockIdx.x * COL_TILE;
  t.ncol = min(COL_TILE, ncols - t.first);
  t.nlev = nlev;
  f(t);
}
#endif

template <class F>
inline int launch_columns(const fv3_ctx *ctx, cudaStream_t st, int i0, int i1, int j0, int j1, int n_arrays, F f) {
  const int ni = i1 - i0, nj = j1 - j0;
  if (ni <= 0 || nj <= 0) return 0;
  const int ncols = ni * nj, ntiles = (ncols + COL_TILE - 1) / COL_TILE, nlev = ctx->g.nz + 1;
  const size_t doubles = (size_t)n_arrays * nlev * COL_TILE;
#ifdef FV3_HOSTSIM
  (void)st;
  const int n_sub = ctx->g.n_sub;
#ifdef FV3_HOSTSIM_OMP
#pragma omp parallel
#endif
  {
    std::vector<double> sm(doubles);
#ifdef FV3_HOSTSIM_OMP
#pragma omp for collapse(2) schedule(static)
#endif
    for (int s = 0; s < n_sub; ++s)
      for (int b = 0; b < ntiles; ++b) {
        Tile t;
        t.sm = sm.data();
        t.s = s;
        t.i0 = i0;
        t.ni = ni;
        t.j0 = j0;
        t.first = b * COL_TILE;
        t.ncol = ncols - t.first < COL_TILE ? ncols - t.first : COL_TILE;
        t.nlev = nlev;
        f(t);
      }
  }
  return 0;
#else
  activate(ctx, st);
  const size_t bytes = doubles * sizeof(double);
  static size_t configured = 0;  // one per template instantiation (= per kernel)
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(kcolumns<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("launch_columns: column tile does not fit in shared memory (too many levels)");
      return (int)e;
    }
    configured = bytes;
  }
  kcolumns<<<dim3(ntiles, ctx->g.n_sub), COL_TILE * COL_WARPS, bytes, st>>>(f, i0, ni, j0, ncols, nlev);
  ++g_launches;
  return 0;
#endif
}

}  // namespace fv3
