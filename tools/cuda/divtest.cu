// Is the shared-reciprocal division of riem_solver.cu bit-identical to a / b?  (gpurun: nvcc -o divtest divtest.cu && ./divtest)
#include <cstdio>
#include <cstdint>
#include <cstring>
struct Recip { double b, r; bool ok; };
__device__ __forceinline__ Recip recip_of(double b) {
  Recip x; x.b = b; double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(b));
  double t = __fma_rn(-b, r0, 1.0); t = __fma_rn(t, t, t);
  double r1 = __fma_rn(r0, t, r0); t = __fma_rn(-b, r1, 1.0); x.r = __fma_rn(r1, t, r1);
  const double ab = fabs(b); x.ok = ab > 1e-290 && ab < 1e290; return x;
}
__device__ __forceinline__ double div_by(double a, const Recip &x) {
  const double aa = fabs(a);
  if (x.ok && aa > 1e-290 && aa < 1e290 && aa < fabs(x.b) * 1e290 && aa * 1e290 > fabs(x.b)) {
    const double q = a * x.r; const double e = __fma_rn(-x.b, q, a); return __fma_rn(x.r, e, q);
  }
  return a / x.b;
}
__device__ uint64_t rng(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
__global__ void k(unsigned long long *bad, unsigned long long *n, int mode) {
  uint64_t s = 88172645463325252ull + (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761ull;
  unsigned long long nb = 0;
  for (int it = 0; it < 20000; ++it) {
    double a, b;
    if (mode == 0) {  // O(1) operands like the solver's
      a = (double)(rng(s) >> 11) * (1.0 / 9007199254740992.0) * 8.0 - 4.0;
      b = (double)(rng(s) >> 11) * (1.0 / 9007199254740992.0) * 6.0 + 0.5;
    } else {          // random bit patterns with moderate exponents
      uint64_t ua = (rng(s) & 0x800fffffffffffffull) | ((uint64_t)(1023 - 40 + (rng(s) % 80)) << 52);
      uint64_t ub = (rng(s) & 0x800fffffffffffffull) | ((uint64_t)(1023 - 40 + (rng(s) % 80)) << 52);
      memcpy(&a, &ua, 8); memcpy(&b, &ub, 8);
    }
    const double q1 = a / b, q2 = div_by(a, recip_of(b));
    if (q1 != q2) ++nb;
  }
  atomicAdd(bad, nb); atomicAdd(n, 20000ull);
}
int main() {
  unsigned long long *d; cudaMalloc(&d, 16);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(d, 0, 16);
    k<<<592, 256>>>(d, d + 1, mode);
    unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode %d: %llu of %llu quotients differ from a / b\n", mode, h[0], h[1]);
  }
  return 0;
}
