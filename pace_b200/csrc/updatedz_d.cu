// Height of the Lagrangian surfaces on the D grid.
//   fv3_update_dz_d <- UpdateHeightOnDGrid.__call__ (fv3core/pace/fv3core/stencils/updatedzd.py:283-356)
#include "common.h"
#include "stream.h"

extern "C" int fv3_fvtp2d(fv3_ctx *, const double *, const double *, const double *, const double *, const double *,
                          double *, double *, const double *, const double *, const double *, int, const double *,
                          const double *, int, int, void *);
extern "C" int fv3_delnflux_nosg(fv3_ctx *, const double *, double *, double *, const double *, const double *, int,
                                 int, void *);

namespace {
constexpr double DZ_MIN = 2.0;
constexpr int NKMAX = 96;
constexpr int SPL_NT = 32;

// One spline column ("chain"): forward elimination with the layer values streamed 8 levels ahead, the partial
// solution in shared memory ([level][lane]) instead of thread-local memory, backward substitution storing the
// interface values.  Same statements, same order as the reference.
struct SplineArgs {
  const double *qc[4];
  double *qi[4];
  const double *gk, *beta, *gamma;
};

template <class COL>
FV_DEV void spline_chain(const fv3_geom &g, const SplineArgs &a, int f, int s, int i, int j, int nz, COL col) {
  const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
  const double *qc = a.qc[f] + c0;
  double *qi = a.qi[f] + c0;
  const double *gk = a.gk, *beta = a.beta, *gamma = a.gamma;
  double qm = FV_LDG(qc), qk = FV_LDG(qc + sk);
  double cm;
  {
    const double xt1 = 2.0 * gk[0] * (gk[0] + 1.0);
    cm = (xt1 * qm + qk) / beta[0];
    col(0) = cm;
  }
  // trip k uses q[k-1], q[k]; the ring delivers q[k+1] for the next trip
  fv3::stream_up(1, nz, [&](int k) { return k + 1 < nz ? FV_LDG(qc + (k + 1) * sk) : 0.0; }, [&](int) { return 0.0; },
                 [&](int k, double q_next, double) {
                   cm = (3.0 * (qm + gk[k] * qk) - cm) / beta[k];
                   col(k) = cm;
                   if (k + 1 < nz) {
                     qm = qk;
                     qk = q_next;
                   }
                 });
  {
    // qk = q[nz-1], qm = q[nz-2], cm = col[nz-1]
    const double gl = gk[nz - 1];
    const double a_bot = 1.0 + gl * (gl + 1.5);
    const double xt1 = 2.0 * gl * (gl + 1.0);
    const double xt2 = gl * (gl + 0.5) - a_bot * gamma[nz - 1];
    cm = (xt1 * qk + qm - a_bot * cm) / xt2;
  }
  qi[nz * sk] = cm;
  for (int k = nz - 1; k >= 0; --k) {
    cm = col(k) - gamma[k] * cm;
    qi[k * sk] = cm;
  }
}

#ifndef FV3_HOSTSIM
__global__ void __launch_bounds__(SPL_NT) kspline(const SplineArgs a, int nz) {
  extern __shared__ double spl_smem[];
  const fv3_geom &g = c_g;
  const int f = (int)blockIdx.x, s = (int)blockIdx.z;
  const int ni = g.nx + 2 * g.halo, nj = g.ny + 2 * g.halo;
  const int idx = (int)blockIdx.y * SPL_NT + (int)threadIdx.x;
  if (idx >= ni * nj) return;
  const int j = idx / ni, i = idx - j * ni;
  double *cs = spl_smem + threadIdx.x;
  spline_chain(g, a, f, s, i, j, nz, [&](int k) -> double & { return cs[k * SPL_NT]; });
}
#endif
}  // namespace

extern "C" {

int fv3_update_dz_d(fv3_ctx *ctx, const double *surface_height, double *height, const double *crx, const double *cry,
                    const double *xfx, const double *yfx, double *ws, double dt, const double *gk, const double *beta,
                    const double *gamma, const double *damp_col, const double *nord_col, int nmax, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  cudaStream_t st = (cudaStream_t)stream;
  const int h = g.halo, nz = g.nz, sj = g.sj;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1, ied = iec + h, jed = jec + h;
  if (nz + 1 > NKMAX) {
    fv3::set_error("fv3_update_dz_d: nz too large");
    return -1;
  }
  double *crx_i = fv3::scratch_field(ctx, 16), *cry_i = fv3::scratch_field(ctx, 17);
  double *xfx_i = fv3::scratch_field(ctx, 18), *yfx_i = fv3::scratch_field(ctx, 19);
  double *fx = fv3::scratch_field(ctx, 20), *fy = fv3::scratch_field(ctx, 21);
  double *gx = fv3::scratch_field(ctx, 22), *gy = fv3::scratch_field(ctx, 23);
  // cubic_spline_interpolation_from_layer_center_to_interfaces (updatedzd.py:157-196), 4 fields, full domain
  SplineArgs sa{{crx, xfx, cry, yfx}, {crx_i, xfx_i, cry_i, yfx_i}, gk, beta, gamma};
#ifdef FV3_HOSTSIM
#ifdef FV3_HOSTSIM_OMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
  for (int s = 0; s < g.n_sub; ++s)
    for (int f = 0; f < 4; ++f)
      for (int j = 0; j <= jed; ++j)
        for (int i = 0; i <= ied; ++i) {
          double col[NKMAX];
          spline_chain(g, sa, f, s, i, j, nz, [&](int k) -> double & { return col[k]; });
        }
#else
  fv3::activate(ctx, st);
  {
    const int ncols = (ied + 1) * (jed + 1);
    kspline<<<dim3(4, (ncols + SPL_NT - 1) / SPL_NT, g.n_sub), SPL_NT, (size_t)(nz + 1) * SPL_NT * sizeof(double), st>>>(sa, nz);
    ++fv3::g_launches;
  }
#endif
  int rc;
  if ((rc = fv3_fvtp2d(ctx, height, crx_i, cry_i, xfx_i, yfx_i, fx, fy, nullptr, nullptr, nullptr, ctx->c.hord_tm, nullptr,
                       nullptr, 0, nz + 1, stream)))
    return rc;
  if ((rc = fv3_delnflux_nosg(ctx, height, gx, gy, damp_col, nord_col, nmax, nz + 1, stream))) return rc;
  // apply_height_fluxes (updatedzd.py:70-126): compute domain, BACKWARD monotonicity fix
  fv3::launch2d(ctx, st, isc, iec + 1, jsc, jec + 1, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    const double ar = m.area[O2(s, i, j)];
    double below = 0.0;
    for (int k = nz; k >= 0; --k) {
      const int64_t o = c0 + k * sk;
      const double area_after = ((ar + xfx_i[o] - xfx_i[o + 1]) + (ar + yfx_i[o] - yfx_i[o + sj])) - ar;
      double hv = (height[o] * ar + fx[o] - fx[o + 1] + fy[o] - fy[o + sj]) / area_after +
                  (gx[o] - gx[o + 1] + gy[o] - gy[o + sj]) / ar;
      if (k == nz) {
        ws[O2(s, i, j)] = (surface_height[O2(s, i, j)] - hv) / dt;
      } else {
        const double other = below + DZ_MIN;
        hv = hv > other ? hv : other;
      }
      height[o] = hv;
      below = hv;
    }
  });
  return fv3::check_launch("fv3_update_dz_d");
}

}  // extern "C"
