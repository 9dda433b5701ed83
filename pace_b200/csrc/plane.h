// Plane-resident kernels: a CTA owns a strip of rows [ja, jb) of one horizontal (subdomain, level) plane, stages it
// — with the 3 halo rows either side that the widest sweep needs — in shared memory and runs a whole multi-sweep
// stage on it as a sequence of block-wide phases.  Every input row is read from HBM once per strip, intermediates
// (inner-sweep fluxes, transversely advected fields, edge values) never leave the SM, and no thread recomputes a
// neighbour's value.  The number of strips per plane is the smallest that lets TWO CTAs share an SM (each CTA's
// barriers and global-load waits are covered by the other's arithmetic); at C128 layout (2,2) that is 2 strips of 32
// rows (39 resident rows of 72 doubles = 22 KB per field).  Rows shared by two strips are computed by both (6 of 38
// rows of the inner sweeps); every output row is stored by exactly one strip.
//
//   launch_planes(ctx, st, k0, k1, n_planes, f)   f(s, k, Block&) runs once per (plane, strip)
//   Block::plane(n)                               n-th shared plane, addressed [j * sj + i] like a global plane
//   Block::rect(i0, i1, j0, j1, g)                g(i, j) over the rectangle, then a block barrier
//   Block::lo(j0, e) / hi(j1, e)                  row range [j0, j1) clipped to [ja - e, jb + e)
//
// Under -DFV3_HOSTSIM (g++, tests only) a "block" is a plain loop, so the same phase code runs on the CPU.
#pragma once
#include <cstdlib>
#include <vector>

#include "common.h"

namespace fv3 {

constexpr int PLANE_THREADS = 512;
// guard doubles before the first / after the last shared plane: the register-window loads of the PPM sweeps (sweep.h)
// run unconditionally and may reach 4 doubles before a row and 2 rows beyond the resident range (never used, never stored)
constexpr int PLANE_PAD_FRONT = 16;
#ifndef FV3_SWEEP_R
#define FV3_SWEEP_R 4
#endif
FV_HD int plane_pad_back(int sj) { return (FV3_SWEEP_R > 4 ? FV3_SWEEP_R - 2 : 2) * sj + 16; }
constexpr int PLANE_SMEM_BUDGET = (233472 / 2) - 1024;  // bytes per CTA for two CTAs per SM (228 KB, 1 KB reserved each)

// t / w for the index decode of a block-wide pass without an integer division (20+ instructions per point): exact for
// t < 2^20, w < 2^10 (every plane that fits in shared memory); `inv` = 1.0f / w
FV_HD int row_of(int t, int w, float inv) {
#ifdef FV3_HOSTSIM
  (void)inv;
  return t / w;
#else
  (void)w;
  return (int)(((float)t + 0.5f) * inv);
#endif
}

struct Block {
  double *sm;
  int pl, off;  // doubles per shared plane; r0 * sj
  int ja, jb;   // owned rows: results on cell rows [ja, jb), y-faces / corner rows [ja, jb]
  int r0, r1;   // resident rows [r0, r1)
  bool last;    // jb is the top of the compute domain (this strip also stores face / corner row jb)
  bool first;   // ja is the bottom of the compute domain
  int s_next, r0_next, nrows_next;  // the (subdomain, strip rows) a CTA of the NEXT wave will start on; s_next < 0: none
  unsigned long long *bar;  // mbarrier of the bulk (TMA) plane loads (device only)
  mutable unsigned phase;   // its current phase parity

  FV_HD double *plane(int n) const { return sm + (int64_t)n * pl - off; }
  FV_HD int jtop() const { return last ? jb : jb - 1; }  // last face / corner row this strip stores
  FV_HD int lo(int j0, int e) const { return j0 > ja - e ? j0 : ja - e; }
  FV_HD int hi(int j1, int e) const { return j1 < jb + e ? j1 : jb + e; }
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
#else
  // One bulk L2 prefetch (cp.async.bulk.prefetch.L2) of n contiguous doubles — the resident rows of a field are
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
#endif
  // f(i, j) for i in [i0, i1), j in [j0, j1)
  template <class F>
  FV_DEV void rect(int i0, int i1, int j0, int j1, F f) const {
    par2(i1 > i0 ? i1 - i0 : 0, j1 > j0 ? j1 - j0 : 0, [&](int ir, int jr) { f(i0 + ir, j0 + jr); });
  }
  // ---- asynchronous plane staging: the resident rows [r0, r1) of a global (s, k) plane are ONE contiguous range, so a
  // whole strip of a field is moved HBM/L2 -> shared memory by a single cp.async.bulk (TMA, no register staging, no
  // per-point address arithmetic); completion is signalled on the CTA's mbarrier.
  //   bulk_begin(n)           one thread arms the barrier for n planes          (block-uniform call)
  //   bulk_rows(dst, plane0)  one thread issues the copy of rows [r0, r1) of the global plane into the shared plane
  //   bulk_wait()             everybody waits for the bytes to land
  FV_HD int bulk_doubles(int sj) const { return (r1 - r0) * sj; }
#ifdef FV3_HOSTSIM
  void bulk_begin(int, int) const {}
  void bulk_rows(double *dst, const double *plane0, int sj) const {
    for (int t = r0 * sj; t < r1 * sj; ++t) dst[t] = plane0[t];
  }
  void bulk_wait() const {}
#else
  __device__ __forceinline__ void bulk_begin(int n_planes, int sj) const {
    __syncthreads();  // every earlier (generic-proxy) access of the destination planes is done
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(n_planes * bulk_doubles(sj) * 8) : "memory");
    }
  }
  __device__ __forceinline__ void bulk_rows(double *dst, const double *plane0, int sj) const {
    if (threadIdx.x == 0) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + r0 * sj), a = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                   "l"(plane0 + r0 * sj), "r"(bulk_doubles(sj) * 8), "r"(a)
                   : "memory");
    }
  }
  __device__ __forceinline__ void bulk_wait() const {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(a), "r"(phase)
          : "memory");
    }
    phase ^= 1u;
  }
#endif
  // resident rows of the global plane starting at `plane0`
  FV_DEV void prefetch_rows(const double *plane0, int sj) const { prefetch_l2(plane0 + r0 * sj, (r1 - r0) * sj); }
  // First-touch operand of a CTA one wave ahead (same level, `field` = start of the 3-D field): its HBM -> L2 transfer
  // starts now, so that CTA's opening load phase finds it in L2.  DRAM is < 15 % busy in these kernels.
  FV_DEV void prefetch_next_wave(const double *field, const fv3_geom &g, int k) const {
    if (s_next >= 0) prefetch_l2(field + (int64_t)s_next * g.ss + (int64_t)k * g.sk + (int64_t)r0_next * g.sj, nrows_next * g.sj);
  }
};

struct StripGeom {
  int ns, rows_per_strip, res_rows;  // strips per plane, owned rows per strip, resident rows per strip
};

// strips per plane for a kernel with n_planes shared planes that runs nk levels; ns == 0: does not fit.
// The smallest number of strips that lets two CTAs share an SM — unless one strip more fills the waves of a SMALL grid
// better: with few subdomains per GPU (4 or 8 GPUs) nk * n_sub * ns CTAs are only a few waves over the 2 * SMs resident
// slots, and the cost of the launch is ~ ceil(waves) * (resident rows per CTA).  Every caller of one stage must pass the
// same (n_planes, nk): the parked-row bookkeeping of the in-place updates depends on the geometry.
inline StripGeom strip_geometry(const fv3_geom &g, int n_planes, int nk = 0) {
  static int forced = -1;
  if (forced < 0) {
    const char *e = getenv("FV3_FORCE_STRIPS");  // tests: exercise the strip logic on small subdomains
    forced = e ? atoi(e) : 0;
  }
  auto geom_of = [&](int ns, StripGeom &sg) {
    const int r = (g.ny + ns - 1) / ns, res = (r + 2 * g.halo + 1 < g.nj) ? r + 2 * g.halo + 1 : g.nj;
    if (r < 4) return -1;
    const int64_t bytes = ((int64_t)n_planes * res * g.sj + PLANE_PAD_FRONT + plane_pad_back(g.sj)) * 8;
    sg = StripGeom{(g.ny + r - 1) / r, r, res};
    return (forced > 0 || bytes <= PLANE_SMEM_BUDGET) ? 1 : 0;
  };
  StripGeom sg{0, 0, 0};
  for (int ns = forced > 0 ? forced : 1; ns <= g.ny; ++ns) {
    StripGeom c;
    const int fit = geom_of(ns, c);
    if (fit < 0) break;
    if (fit == 0) continue;
    sg = c;
    if (forced <= 0 && nk > 0) {
      // one strip more, if it fits and costs fewer (rounded-up waves) x (rows per CTA)
      static int slots = 0;
      if (slots == 0) {
#ifdef FV3_HOSTSIM
        slots = 296;
#else
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = 2 * sms;
#endif
      }
      StripGeom d;
      if (geom_of(ns + 1, d) == 1) {
        auto cost = [&](const StripGeom &x) {
          const int64_t ctas = (int64_t)nk * g.n_sub * x.ns;
          return ((ctas + slots - 1) / slots) * x.res_rows;
        };
        if (cost(d) * 100 < cost(sg) * 97) sg = d;  // at least 3 % better by the model
      }
    }
    break;
  }
  return sg;
}

FV_HD void strip_rows(const fv3_geom &g, int strip, int rows_per_strip, int &ja, int &jb, int &r0, int &r1) {
  const int jend = g.halo + g.ny;
  ja = g.halo + strip * rows_per_strip;
  jb = ja + rows_per_strip < jend ? ja + rows_per_strip : jend;
  r0 = ja - g.halo;
  r1 = jb + g.halo + 1 < g.nj ? jb + g.halo + 1 : g.nj;
}

FV_HD Block make_block(const fv3_geom &g, double *sm, int strip, int rows_per_strip, int res_rows) {
  Block b;
  b.s_next = -1;
  b.r0_next = b.nrows_next = 0;
  b.sm = sm + PLANE_PAD_FRONT;
  const int jsc = g.halo, jend = g.halo + g.ny;
  b.ja = jsc + strip * rows_per_strip;
  b.jb = b.ja + rows_per_strip < jend ? b.ja + rows_per_strip : jend;
  b.last = b.jb == jend;
  b.first = strip == 0;
  b.bar = nullptr;
  b.phase = 0;
  b.r0 = b.ja - g.halo;
  b.r1 = b.jb + g.halo + 1 < g.nj ? b.jb + g.halo + 1 : g.nj;
  b.pl = res_rows * g.sj;
  b.off = b.r0 * g.sj;
  return b;
}

#ifndef FV3_HOSTSIM
template <class F>
__global__ void __launch_bounds__(PLANE_THREADS, 2) kplane(F f, int k0, int rows_per_strip, int res_rows, int ahead) {
  extern __shared__ __align__(128) double plane_smem[];
  __shared__ unsigned long long plane_bar;
  Block b = make_block(c_g, plane_smem, (int)blockIdx.z, rows_per_strip, res_rows);
  b.bar = &plane_bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&plane_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  {  // blocks are dispatched level-fastest, then subdomain, then strip: `ahead` subdomains later = one wave later
    int sn = (int)blockIdx.y + ahead, zn = (int)blockIdx.z;
    if (sn >= (int)gridDim.y) {
      sn -= (int)gridDim.y;
      ++zn;
    }
    if (zn < (int)gridDim.z && sn < (int)gridDim.y) {
      int ja, jb, r0, r1;
      strip_rows(c_g, zn, rows_per_strip, ja, jb, r0, r1);
      b.s_next = sn;
      b.r0_next = r0;
      b.nrows_next = r1 - r0;
    }
  }
  f((int)blockIdx.y, k0 + (int)blockIdx.x, b);
}
#endif

template <class F>
inline int launch_planes(const fv3_ctx *ctx, cudaStream_t st, int k0, int k1, int n_planes, F f) {
  if (k1 <= k0) return 0;
  const StripGeom sg = strip_geometry(ctx->g, n_planes, k1 - k0);
  if (sg.ns == 0) {
    set_error("launch_planes: no strip decomposition of the plane fits in shared memory");
    return -1;
  }
  const size_t doubles = (size_t)n_planes * sg.res_rows * ctx->g.sj + PLANE_PAD_FRONT + plane_pad_back(ctx->g.sj);
#ifdef FV3_HOSTSIM
  (void)st;
  const int n_sub = ctx->g.n_sub;
  const fv3_geom gg = ctx->g;
#ifdef FV3_HOSTSIM_OMP
#pragma omp parallel
#endif
  {
    std::vector<double> sm(doubles);
#ifdef FV3_HOSTSIM_OMP
#pragma omp for collapse(2) schedule(static)
#endif
    for (int s = 0; s < n_sub; ++s)
      for (int k = k0; k < k1; ++k)
        for (int z = 0; z < sg.ns; ++z) {
          const Block b = make_block(gg, sm.data(), z, sg.rows_per_strip, sg.res_rows);
          f(s, k, b);
        }
  }
  return 0;
#else
  activate(ctx, st);
  const size_t bytes = doubles * sizeof(double);
  static size_t configured = 0;  // one per template instantiation (= per kernel)
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(kplane<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("launch_planes: strip does not fit in shared memory");
      return (int)e;
    }
    configured = bytes;
  }
  static int resident = 0;  // CTAs of this kernel resident on the device at once (2 per SM)
  if (resident == 0) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    resident = 2 * sms;
  }
  const int ahead = (resident + (k1 - k0) - 1) / (k1 - k0);
  kplane<<<dim3(k1 - k0, ctx->g.n_sub, sg.ns), PLANE_THREADS, bytes, st>>>(f, k0, sg.rows_per_strip, sg.res_rows, ahead);
  ++g_launches;
  return 0;
#endif
}

}  // namespace fv3
