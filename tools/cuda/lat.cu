// Dependent-issue latencies of the fp64 operations on the critical path of the column recurrences (one warp alone on an
// SM), and cycles per level of the Thomas forward sweep as written in riem_solver.cu.   nvcc -arch=sm_100a -fmad=false
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../../pace_b200/csrc/fdiv.h"
using fv3::div_by;
using fv3::Recip;
using fv3::recip_of;

__global__ void lat(double *out, long long *cyc, double a, double b) {
  double x = a;
  long long t0, t1;
  const int N = 4096;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __fma_rn(x, b, a);
  t1 = clock64();
  cyc[0] = t1 - t0;
  double y = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) y = __dadd_rn(y, a);
  t1 = clock64();
  cyc[1] = t1 - t0;
  double z = y;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) z = __dmul_rn(z, b);
  t1 = clock64();
  cyc[2] = t1 - t0;
  double r = z + 1.5;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r));
  t1 = clock64();
  cyc[3] = t1 - t0;
  double q = r + 2.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) q = a / (q + b);   // compiler's full division + add
  t1 = clock64();
  cyc[4] = t1 - t0;
  // select chain: dmin
  double m = q;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) m = (m < a ? m : a) + b;
  t1 = clock64();
  cyc[5] = t1 - t0;
  out[threadIdx.x] = x + y + z + r + q + m;
}

// the pp forward elimination of sim1_tile on a [level][32] shared array
__global__ void thomas(double *out, long long *cyc, int nz, double seed) {
  extern __shared__ double sm_[];
  double *A = sm_, *B = sm_ + 80 * 32, *C = sm_ + 160 * 32;
  const int c = threadIdx.x, T = 32;
  for (int k = 0; k <= nz; ++k) {
    C[k * T + c] = 1000.0 + seed * (k + c);
    if (k < nz) B[k * T + c] = 1.0 + 0.01 * seed * k;
  }
  __syncthreads();
  long long t0 = clock64();
  {
    constexpr int UNR = 4;
    double pe_k = C[c], pe_n = C[T + c], gr = B[c];
    double bet = 2.0 * (1.0 + gr);
    Recip rb = recip_of(bet);
    double pp = div_by(3.0 * (pe_k + gr * pe_n), rb);
    C[c] = 0.0;
    C[T + c] = pp;
    int k = 1;
    for (; k + UNR <= nz - 1; k += UNR) {
      double pn[UNR], g2[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        pn[u] = C[(k + u + 1) * T + c];
        g2[u] = B[(k + u) * T + c];
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const double gam = div_by(gr, rb);
        pe_k = pe_n;
        pe_n = pn[u];
        gr = g2[u];
        const double bb = 2.0 * (1.0 + gr), dd = 3.0 * (pe_k + gr * pe_n);
        bet = bb - gam;
        rb = recip_of(bet);
        pp = div_by(dd - pp, rb);
        A[(k + u) * T + c] = gam;
        C[(k + u + 1) * T + c] = pp;
      }
    }
  }
  long long t1 = clock64();
  cyc[8] = t1 - t0;
  // back substitution
  t0 = clock64();
  {
    double ppn = C[nz * T + c];
    for (int k = nz - 1; k >= 8; k -= 8) {
      double cc[8], aa[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        cc[u] = C[(k - u) * T + c];
        aa[u] = A[(k - u) * T + c];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        ppn = cc[u] - aa[u] * ppn;
        C[(k - u) * T + c] = ppn;
      }
    }
  }
  t1 = clock64();
  cyc[9] = t1 - t0;
  out[c] = C[(nz - 1) * T + c] + A[5 * T + c];
}

// the pp forward elimination of sim1_tile on a [level][32] shared array
__global__ void thomas2(double *out, long long *cyc, int nz, double seed) {
  extern __shared__ double sm_[];
  double *A = sm_, *B = sm_ + 80 * 32, *C = sm_ + 160 * 32;
  const int c = threadIdx.x, T = 32;
  for (int k = 0; k <= nz; ++k) {
    C[k * T + c] = 1000.0 + seed * (k + c);
    if (k < nz) B[k * T + c] = 1.0 + 0.01 * seed * k;
  }
  __syncthreads();
  long long t0 = clock64();
  {
    constexpr int UNR = 4;
    double pe_k = C[c], pe_n = C[T + c], gr = B[c];
    double bet = 2.0 * (1.0 + gr);
    Recip rb = recip_of(bet);
    double pp = div_by(3.0 * (pe_k + gr * pe_n), rb);
    C[c] = 0.0;
    C[T + c] = pp;
    int k = 1;
    for (; k + UNR <= nz - 1; k += UNR) {
      double pn[UNR], g2[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        pn[u] = C[(k + u + 1) * T + c];
        g2[u] = B[(k + u) * T + c];
      }
      double gam_o[UNR], pp_o[UNR];
      const double pe_n0 = pe_n, gr0 = gr, bet0 = bet, pp0 = pp;
      bool bad = !rb.ok;
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const double gam = fv3::div_fast(gr, rb, bad);
        pe_k = pe_n;
        pe_n = pn[u];
        gr = g2[u];
        const double bb = 2.0 * (1.0 + gr), dd = 3.0 * (pe_k + gr * pe_n);
        bet = bb - gam;
        rb = fv3::recip_fast(bet, bad);
        pp = fv3::div_fast(dd - pp, rb, bad);
        gam_o[u] = gam;
        pp_o[u] = pp;
      }
      if (bad) {
        pe_n = pe_n0; gr = gr0; bet = bet0; pp = pp0;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const double gam = gr / bet;
          pe_k = pe_n;
          pe_n = pn[u];
          gr = g2[u];
          const double bb = 2.0 * (1.0 + gr), dd = 3.0 * (pe_k + gr * pe_n);
          bet = bb - gam;
          pp = (dd - pp) / bet;
          gam_o[u] = gam;
          pp_o[u] = pp;
        }
        rb = recip_of(bet);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        A[(k + u) * T + c] = gam_o[u];
        C[(k + u + 1) * T + c] = pp_o[u];
      }
    }
  }
  long long t1 = clock64();
  cyc[10] = t1 - t0;
  // back substitution
  t0 = clock64();
  {
    double ppn = C[nz * T + c];
    for (int k = nz - 1; k >= 8; k -= 8) {
      double cc[8], aa[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        cc[u] = C[(k - u) * T + c];
        aa[u] = A[(k - u) * T + c];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        ppn = cc[u] - aa[u] * ppn;
        C[(k - u) * T + c] = ppn;
      }
    }
  }
  t1 = clock64();
  cyc[11] = t1 - t0;
  out[c] = C[(nz - 1) * T + c] + A[5 * T + c];
}

int main() {
  double *out;
  long long *cyc, h[16];
  double ha[32], hb[32];
  cudaMalloc(&out, 1024);
  cudaMalloc(&cyc, 128);
  for (int rep = 0; rep < 2; ++rep) {
    lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    cudaFuncSetAttribute(thomas, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    thomas<<<1, 32, 62 * 1024>>>(out, cyc, 79, 0.37);
    cudaMemcpy(ha, out, 256, cudaMemcpyDeviceToHost);
    cudaFuncSetAttribute(thomas2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    thomas2<<<1, 32, 62 * 1024>>>(out, cyc, 79, 0.37);
    cudaMemcpy(hb, out, 256, cudaMemcpyDeviceToHost);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, cyc, 128, cudaMemcpyDeviceToHost);
  printf("dependent latency (cycles/op): DFMA %.1f  DADD %.1f  DMUL %.1f  RCP64H %.1f  div+add %.1f  min+add %.1f\n", h[0] / 4096.0,
         h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0, h[4] / 4096.0, h[5] / 4096.0);
  printf("thomas forward: %lld cycles = %.1f per level; back substitution: %lld cycles = %.1f per level\n", h[8], h[8] / 76.0, h[9],
         h[9] / 72.0);
  printf("branch-free trips: %lld cycles = %.1f per level; results %s\n", h[10], h[10] / 76.0, memcmp(ha, hb, 256) ? "DIFFER" : "identical");
  return 0;
}
