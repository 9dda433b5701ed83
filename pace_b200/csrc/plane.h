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
  void prefetch_l2(const double *, int) const {}
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
  // as par2, with the global-memory operands of a point fetched by `load(ir, jr, v)` into NV registers before
  // `comp(ir, jr, v)` runs
  template <int NV, class L, class C>
  void par2_pre(int w, int nrows, L load, C comp) const {
    for (int jr = 0; jr < nrows; ++jr)
      for (int ir = 0; ir < w; ++ir) {
        double v[NV];
        load(ir, jr, v);
        comp(ir, jr, v);
      }
  }
#else
  // One bulk L2 prefetch (cp.async.bulk.prefetch.L2) of n contiguous doubles — a whole (s, k) plane of a field is
  // contiguous in HBM — issued by one thread: operands of later phases are pulled into L2 while the CTA computes.
  __device__ __forceinline__ void prefetch_l2(const double *p, int n) const {
    if (threadIdx.x == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((n * 8) & ~15) : "memory");
  }
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
  // Software-pipelined form: each thread first issues the global loads of 2 or 4 of its points (independent
  // LDG.NC requests in flight; out-of-range slots re-load the first point, results unused), then computes them.  With one CTA per SM (shared memory bound) this is what hides
  // the L2 / HBM latency that the barrier-separated phases would otherwise expose.
  template <int NV, class L, class C>
  __device__ __forceinline__ void par2_pre(int w, int nrows, L load, C comp) const {
    const int n = w * nrows, nt = blockDim.x;
    const float inv = 1.0f / (float)w;
    if (NV >= 4) {
      for (int t0 = threadIdx.x; t0 < n; t0 += 2 * nt) {
        const int t1 = t0 + nt;
        const bool h1 = t1 < n;
        const int j0 = (int)(((float)t0 + 0.5f) * inv), i0 = t0 - j0 * w;
        const int j1 = h1 ? (int)(((float)t1 + 0.5f) * inv) : j0, i1 = h1 ? t1 - j1 * w : i0;
        double v0[NV], v1[NV];
        load(i0, j0, v0);
        load(i1, j1, v1);
        comp(i0, j0, v0);
        if (h1) comp(i1, j1, v1);
      }
    } else {
      for (int t0 = threadIdx.x; t0 < n; t0 += 4 * nt) {
        const int t1 = t0 + nt, t2 = t1 + nt, t3 = t2 + nt;
        const bool h1 = t1 < n, h2 = t2 < n, h3 = t3 < n;
        const int j0 = (int)(((float)t0 + 0.5f) * inv), i0 = t0 - j0 * w;
        const int j1 = h1 ? (int)(((float)t1 + 0.5f) * inv) : j0, i1 = h1 ? t1 - j1 * w : i0;
        const int j2 = h2 ? (int)(((float)t2 + 0.5f) * inv) : j0, i2 = h2 ? t2 - j2 * w : i0;
        const int j3 = h3 ? (int)(((float)t3 + 0.5f) * inv) : j0, i3 = h3 ? t3 - j3 * w : i0;
        double v0[NV], v1[NV], v2[NV], v3[NV];
        load(i0, j0, v0);
        load(i1, j1, v1);
        load(i2, j2, v2);
        load(i3, j3, v3);
        comp(i0, j0, v0);
        if (h1) comp(i1, j1, v1);
        if (h2) comp(i2, j2, v2);
        if (h3) comp(i3, j3, v3);
      }
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
