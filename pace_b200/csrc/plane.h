// Plane-resident kernels: one CTA owns one horizontal (subdomain, level) plane — (nx+7) x (ny+7) points, 40 KB per
// fp64 field at C128 layout (2,2) — stages it in shared memory and runs a whole multi-sweep stage on it as a
// sequence of block-wide phases.  Every input plane is read from HBM once, intermediates (inner-sweep fluxes,
// transversely advected fields, edge values) never leave the SM, and no thread recomputes a neighbour's value.
//
//   launch_planes(ctx, st, k0, k1, smem_doubles, f)   f(s, k, Block&) runs once per plane
//   Block::par(n, g)                                   g(t) for t in [0, n), then a block barrier
//
// Under -DFV3_HOSTSIM (g++, tests only) a "block" is a plain loop, so the same phase code runs on the CPU.
#pragma once
#include <vector>

#include "common.h"

namespace fv3 {

constexpr int PLANE_THREADS = 1024;

struct Block {
  double *sm;
#ifdef FV3_HOSTSIM
  template <class F>
  void par(int n, F f) const {
    for (int t = 0; t < n; ++t) f(t);
  }
  // f(ir, jr) for ir in [0, w), jr in [0, nrows)
  template <class F>
  void par2(int w, int nrows, F f) const {
    for (int jr = 0; jr < nrows; ++jr)
      for (int ir = 0; ir < w; ++ir) f(ir, jr);
  }
#else
  template <class F>
  __device__ __forceinline__ void par(int n, F f) const {
    for (int t = threadIdx.x; t < n; t += blockDim.x) f(t);
    __syncthreads();
  }
  // row-major (ir fastest) over w x nrows points; the row index comes from a float reciprocal (exact for the
  // plane sizes that fit in shared memory: t < 2^20, w < 2^10) instead of an integer division per point
  template <class F>
  __device__ __forceinline__ void par2(int w, int nrows, F f) const {
    const int n = w * nrows;
    const float inv = 1.0f / (float)w;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      const int jr = (int)(((float)t + 0.5f) * inv);
      f(t - jr * w, jr);
    }
    __syncthreads();
  }
#endif
};

#ifndef FV3_HOSTSIM
template <class F>
__global__ void __launch_bounds__(PLANE_THREADS) kplane(F f, int k0) {
  extern __shared__ double plane_smem[];
  Block b{plane_smem};
  f((int)blockIdx.y, k0 + (int)blockIdx.x, b);
}
#endif

template <class F>
inline int launch_planes(const fv3_ctx *ctx, cudaStream_t st, int k0, int k1, int smem_doubles, F f) {
  if (k1 <= k0) return 0;
#ifdef FV3_HOSTSIM
  (void)st;
  const int n_sub = ctx->g.n_sub;
#ifdef FV3_HOSTSIM_OMP
#pragma omp parallel
#endif
  {
    std::vector<double> sm((size_t)smem_doubles);
    Block b{sm.data()};
#ifdef FV3_HOSTSIM_OMP
#pragma omp for collapse(2) schedule(static)
#endif
    for (int s = 0; s < n_sub; ++s)
      for (int k = k0; k < k1; ++k) f(s, k, b);
  }
  return 0;
#else
  activate(ctx, st);
  const size_t bytes = (size_t)smem_doubles * sizeof(double);
  static size_t configured = 0;  // one per template instantiation (= per kernel)
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(kplane<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("launch_planes: plane does not fit in shared memory (subdomain too large for a plane-resident kernel)");
      return (int)e;
    }
    configured = bytes;
  }
  kplane<<<dim3(k1 - k0, ctx->g.n_sub), PLANE_THREADS, bytes, st>>>(f, k0);
  ++g_launches;
  return 0;
#endif
}

}  // namespace fv3
