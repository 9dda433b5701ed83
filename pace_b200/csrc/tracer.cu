// Sub-cycled tracer advection.
//   fv3_tracer_flux_prep      <- flux_compute + divide_fluxes_by_n_substeps (tracer_2d_1l.py:17-100)
//   fv3_tracer_apply_mass_flux<- apply_mass_flux (:115-135)
//   fv3_tracer_apply_flux     <- apply_tracer_flux (:138-156)
//   fv3_tracer_swap_dp        <- swap_dp (:159-163)
// The 2-D transport itself is fv3_fvtp2d (hord_tr).
#include "common.h"

extern "C" {

int fv3_tracer_flux_prep(fv3_ctx *ctx, double *cxd, double *cyd, double *mfxd, double *mfyd, double *xfx, double *yfx,
                         int n_split, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo, sj = g.sj;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  const double frac = 1.0 / n_split;
  fv3::launch3d(ctx, (cudaStream_t)stream, 0, g.ni, 0, g.nj, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    if (i >= isc && i <= iec + 1 && j >= jsc - 3 && j <= jec + 3) {
      const double cx = cxd[o];
      xfx[o] = cx > 0 ? cx * m.dxa[o2 - 1] * m.dy[o2] * m.sin_sg3[o2 - 1] : cx * m.dxa[o2] * m.dy[o2] * m.sin_sg1[o2];
    }
    if (i >= isc - 3 && i <= iec + 3 && j >= jsc && j <= jec + 1) {
      const double cy = cyd[o];
      yfx[o] = cy > 0 ? cy * m.dya[o2 - sj] * m.dx[o2] * m.sin_sg4[o2 - sj] : cy * m.dya[o2] * m.dx[o2] * m.sin_sg2[o2];
    }
    if (n_split > 1) {
      cxd[o] = cxd[o] * frac;
      xfx[o] = xfx[o] * frac;
      mfxd[o] = mfxd[o] * frac;
      cyd[o] = cyd[o] * frac;
      yfx[o] = yfx[o] * frac;
      mfyd[o] = mfyd[o] * frac;
    }
  });
  return fv3::check_launch("fv3_tracer_flux_prep");
}

int fv3_tracer_apply_mass_flux(fv3_ctx *ctx, const double *dp1, const double *mfx, const double *mfy, double *dp2,
                               void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo, sj = g.sj;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    dp2[o] = dp1[o] + (mfx[o] - mfx[o + 1] + mfy[o] - mfy[o + sj]) * m.rarea[O2(s, i, j)];
  });
  return fv3::check_launch("fv3_tracer_apply_mass_flux");
}

int fv3_tracer_apply_flux(fv3_ctx *ctx, double *q, const double *dp1, const double *fx, const double *fy,
                          const double *dp2, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo, sj = g.sj;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    q[o] = (q[o] * dp1[o] + (fx[o] - fx[o + 1] + fy[o] - fy[o + sj]) * m.rarea[O2(s, i, j)]) / dp2[o];
  });
  return fv3::check_launch("fv3_tracer_apply_flux");
}

int fv3_tracer_swap_dp(fv3_ctx *ctx, double *dp1, double *dp2, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    const double t = dp1[o];
    dp1[o] = dp2[o];
    dp2[o] = t;
  });
  return fv3::check_launch("fv3_tracer_swap_dp");
}

}  // extern "C"

extern "C" {

// CubedToLatLon ord4_transform (stencils/pace/stencils/c2l_ord.py:41-66) on the compute domain
int fv3_c2l_ord4(fv3_ctx *ctx, const double *u, const double *v, double *ua, double *va, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo, sj = g.sj;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  const double C1 = 1.125, C2 = -0.125;
  fv3::launch3d(ctx, (cudaStream_t)stream, isc, iec + 1, jsc, jec + 1, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    double utmp, vtmp;
    if ((S && j == jsc) || (N && j == jec) || (W && i == isc) || (E && i == iec)) {
      utmp = 2.0 * (u[o] * m.dx[o2] + u[o + sj] * m.dx[o2 + sj]) / (m.dx[o2] + m.dx[o2 + sj]);
      vtmp = 2.0 * ((v[o] * m.dy[o2]) + (v[o + 1] * m.dy[o2 + 1])) / (m.dy[o2] + m.dy[o2 + 1]);
    } else {
      utmp = C2 * (u[o - sj] + u[o + 2 * sj]) + C1 * (u[o] + u[o + sj]);
      vtmp = C2 * (v[o - 1] + v[o + 2]) + C1 * (v[o] + v[o + 1]);
    }
    ua[o] = m.a11[o2] * utmp + m.a12[o2] * vtmp;
    va[o] = m.a21[o2] * utmp + m.a22[o2] * vtmp;
  });
  return fv3::check_launch("fv3_c2l_ord4");
}

}  // extern "C"
