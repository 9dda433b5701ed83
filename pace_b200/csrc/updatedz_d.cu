// Height of the Lagrangian surfaces on the D grid.
//   fv3_update_dz_d <- UpdateHeightOnDGrid.__call__ (fv3core/pace/fv3core/stencils/updatedzd.py:283-356)
#include "common.h"
#include "fdiv.h"
#include "transport.h"

namespace {
constexpr double DZ_MIN = 2.0;
constexpr int NKMAX = 96;

// Transport of the interface heights with the interpolated Courant numbers / area fluxes (updatedzd.py:331-343), the
// del-n damping fluxes of the heights (DelnFluxNoSG, :344-345) and the flux-form update of apply_height_fluxes
// (:70-106) as ONE strip-resident kernel: the four flux fields stay in shared memory (B, A: transport; D, T: damping)
// and are applied in the CTA that produced them.  The heights are only read here (the update goes to `hraw`), so no
// row needs parking.
template <int MORD>
int height_fluxes_launch(const fv3_ctx *ctx, cudaStream_t st, const double *height, const double *crx_i, const double *cry_i,
                         const double *xfx_i, const double *yfx_i, double *hraw, const double *damp_col,
                         const double *nord_col, int nmax) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  return fv3::launch_planes(ctx, st, 0, g.nz + 1, fv3::FVTP_PLANES, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *Q = b.plane(0), *A = b.plane(1), *B = b.plane(2), *D = b.plane(3), *T = b.plane(4);
    const int sj = g.sj, h = g.halo, nx = g.nx;
    const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
    const fv3::PlaneArgs pa{height, crx_i, cry_i, xfx_i, yfx_i, xfx_i, yfx_i};
    fv3::fvtp2d_plane<MORD>(g, m, s, k, b, pa, Q, A, B, D, T);
    fv3::delnflux_plane(g, m, s, b, height + ob, damp_col[k], nord_col[k] > 0, nmax, false, Q, D, T);
    const double *area = m.area + o2b, *xf = xfx_i + ob, *yf = yfx_i + ob, *hg = height + ob;
    b.rect(h, h + nx, b.ja, b.jb, [&](int i, int j) {
      const int p = j * sj + i;
      const double ar = FV_LDG(area + p);
      const double area_after = ((ar + FV_LDG(xf + p) - FV_LDG(xf + p + 1)) + (ar + FV_LDG(yf + p) - FV_LDG(yf + p + sj))) - ar;
      hraw[ob + p] = (FV_LDG(hg + p) * ar + B[p] - B[p + 1] + A[p] - A[p + sj]) / area_after + (D[p] - D[p + 1] + T[p] - T[p + sj]) / ar;
    });
  });
}
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
  // cubic_spline_interpolation_from_layer_center_to_interfaces (updatedzd.py:157-196), 4 fields, full domain.
  // One thread per column and field, partial solution in thread-local memory: at 472 K independent columns the
  // full-occupancy form beats shared-memory chains (measured: 384 us against 590 us, profiles/).
  fv3::launch3d(ctx, st, 0, ied + 1, 0, jed + 1, 0, 4, FV_LAMBDA(int s, int i, int j, int f) { FV_DEV_GM
    const double *qc = f == 0 ? crx : (f == 1 ? xfx : (f == 2 ? cry : yfx));
    double *qi = f == 0 ? crx_i : (f == 1 ? xfx_i : (f == 2 ? cry_i : yfx_i));
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    double col[NKMAX];
    {
      const double xt1 = 2.0 * gk[0] * (gk[0] + 1.0);
      col[0] = (xt1 * qc[c0] + qc[c0 + sk]) / beta[0];
    }
    // forward elimination: one dependent divide per level.  The reciprocals of beta (the same for every column) are
    // refined four levels at a time AHEAD of the chain, which then costs a multiply and two FMAs per level (fdiv.h)
    int k = 1;
    double prev = col[0];
    for (; k + 3 < nz; k += 4) {
      fv3::Recip rb[4];
      double rhs[4], qv[5];
#pragma unroll
      for (int n = 0; n < 4; ++n) rb[n] = fv3::recip_of(beta[k + n]);
#pragma unroll
      for (int n = 0; n < 5; ++n) qv[n] = qc[c0 + (k - 1 + n) * sk];
#pragma unroll
      for (int n = 0; n < 4; ++n) rhs[n] = 3.0 * (qv[n] + gk[k + n] * qv[n + 1]);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        prev = fv3::div_by(rhs[n] - prev, rb[n]);
        col[k + n] = prev;
      }
    }
    for (; k < nz; ++k) col[k] = (3.0 * (qc[c0 + (k - 1) * sk] + gk[k] * qc[c0 + k * sk]) - col[k - 1]) / beta[k];
    {
      const double gl = gk[nz - 1];
      const double a_bot = 1.0 + gl * (gl + 1.5);
      const double xt1 = 2.0 * gl * (gl + 1.0);
      const double xt2 = gl * (gl + 0.5) - a_bot * gamma[nz - 1];
      col[nz] = (xt1 * qc[c0 + (nz - 1) * sk] + qc[c0 + (nz - 2) * sk] - a_bot * col[nz - 1]) / xt2;
    }
    qi[c0 + nz * sk] = col[nz];
    // back substitution: partial solutions and gammas fetched 8 levels ahead of the multiply-subtract chain
    double above = col[nz];
    int kb = nz - 1;
    for (; kb - 7 >= 0; kb -= 8) {
      double cv[8], gv[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        cv[n] = col[kb - n];
        gv[n] = gamma[kb - n];
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        above = cv[n] - gv[n] * above;
        qi[c0 + (kb - n) * sk] = above;
      }
    }
    for (; kb >= 0; --kb) {
      above = col[kb] - gamma[kb] * above;
      qi[c0 + kb * sk] = above;
    }
  });
  // transport + del-n damping + apply_height_fluxes (updatedzd.py:70-126, 331-345), compute domain, one launch.  The
  // flux-form update of every level is independent (into a scratch field); only the BACKWARD monotonicity fix is a column
  // walk, with its loads fetched 8 levels ahead of the dependent max-chain.
  if (nmax > 2 || nmax < 0) {
    fv3::set_error("fv3_update_dz_d: nmax must be 0..2 (halo 3)");
    return -1;
  }
  double *hraw = fv3::scratch_field(ctx, 24);
  const int mord = ctx->c.hord_tm < 0 ? -ctx->c.hord_tm : ctx->c.hord_tm;
  int rc;
  if (mord == 8 || mord == 10)
    rc = height_fluxes_launch<8>(ctx, st, height, crx_i, cry_i, xfx_i, yfx_i, hraw, damp_col, nord_col, nmax);
  else if (mord == 5)
    rc = height_fluxes_launch<5>(ctx, st, height, crx_i, cry_i, xfx_i, yfx_i, hraw, damp_col, nord_col, nmax);
  else
    rc = height_fluxes_launch<6>(ctx, st, height, crx_i, cry_i, xfx_i, yfx_i, hraw, damp_col, nord_col, nmax);
  if (rc) return rc;
  fv3::launch2d(ctx, st, isc, iec + 1, jsc, jec + 1, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    double below = hraw[c0 + nz * sk];
    ws[O2(s, i, j)] = (surface_height[O2(s, i, j)] - below) / dt;
    height[c0 + nz * sk] = below;
    int k = nz - 1;
    for (; k - 7 >= 0; k -= 8) {
      double v[8];
#pragma unroll
      for (int n = 0; n < 8; ++n) v[n] = hraw[c0 + (k - n) * sk];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const double other = below + DZ_MIN;
        below = v[n] > other ? v[n] : other;
        height[c0 + (k - n) * sk] = below;
      }
    }
    for (; k >= 0; --k) {
      const double hv = hraw[c0 + k * sk], other = below + DZ_MIN;
      below = hv > other ? hv : other;
      height[c0 + k * sk] = below;
    }
  });
  return fv3::check_launch("fv3_update_dz_d");
}

}  // extern "C"
