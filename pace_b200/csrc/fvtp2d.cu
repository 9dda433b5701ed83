// 2-D flux-form finite-volume transport and the del-n diffusive fluxes.
//   fv3_fvtp2d        <- FiniteVolumeTransport.__call__ (fv3core/pace/fv3core/stencils/fvtp2d.py:235-346) incl.
//                        XPiecewiseParabolic / YPiecewiseParabolic (xppm.py:269-353, yppm.py) and, when a damping
//                        column is given, DelnFlux.__call__ (delnflux.py:1164-1207)
//   fv3_delnflux_nosg <- DelnFluxNoSG.__call__ (delnflux.py:1209-1261)
// 10 + (7 + 6*nmax) reference launches become 3 + (2 + nmax): the corner copies are index remaps at read time
// (ppm.h), the Laplacian d2 of each del-n iteration is recomputed from the previous fluxes instead of stored.
#include "common.h"
#include "ppm.h"

namespace {

constexpr int S_FYIN = 8, S_FXIN = 9, S_QI = 10, S_QJ = 11, S_DA = 12;  // scratch slots (12..15 = delnflux ping-pong)

struct Idx {
  int isc, iec, jsc, jec, ied, jed;
};
Idx make_idx(const fv3_geom &g) {
  Idx x;
  x.isc = g.halo;
  x.iec = g.halo + g.nx - 1;
  x.jsc = g.halo;
  x.jec = g.halo + g.ny - 1;
  x.ied = x.iec + g.halo;
  x.jed = x.jec + g.halo;
  return x;
}

// accumulate: 0 = leave fx2/fy2 in (out_fx, out_fy); 1 = out_fx += fx2 ; 2 = out_fx += 0.5*damp*(mass[-1]+mass)*fx2
void delnflux_core(const fv3_ctx *ctx, cudaStream_t st, const double *q, const double *damp, const double *nord,
                   int nmax, int nk, bool copy_q, double *out_fx, double *out_fy, int accumulate, const double *mass) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const Idx x = make_idx(g);
  const int sj = g.sj;
  double *bufx[2] = {fv3::scratch_field(ctx, S_DA), fv3::scratch_field(ctx, S_DA + 2)};
  double *bufy[2] = {fv3::scratch_field(ctx, S_DA + 1), fv3::scratch_field(ctx, S_DA + 3)};
  const int isc = x.isc, iec = x.iec, jsc = x.jsc, jec = x.jec;
  {
    double *fxo = bufx[0], *fyo = bufy[0];
    // d2_damp_interval / copy_stencil_interval + corner copies + fx/fy_calc_stencil_nord (delnflux.py:59-126,...)
    fv3::launch3d(ctx, st, isc - nmax, iec + 2 + nmax, jsc - nmax, jec + 2 + nmax, 0, nk, FV_LAMBDA(int s, int i, int j, int k) {
      const bool hi = nord[k] > 0;
      const int r = hi ? nmax : 0;
      const double dk = copy_q ? 1.0 : damp[k];
      auto d2x = [&](int ii, int jj) {
        if (hi) fv3::corner_x(g, s, ii, jj);
        const double v = q[O3(s, ii, jj, k)];
        return copy_q ? v : dk * v;
      };
      auto d2y = [&](int ii, int jj) {
        if (hi) fv3::corner_y(g, s, ii, jj);
        const double v = q[O3(s, ii, jj, k)];
        return copy_q ? v : dk * v;
      };
      const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
      if (i >= isc - r && i <= iec + 1 + r && j >= jsc - r && j <= jec + r) fxo[o] = m.del6_v[o2] * (d2x(i - 1, j) - d2x(i, j));
      if (i >= isc - r && i <= iec + r && j >= jsc - r && j <= jec + 1 + r) fyo[o] = m.del6_u[o2] * (d2y(i, j - 1) - d2y(i, j));
    });
  }
  int cur = 0;
  for (int n = 0; n < nmax; ++n) {
    const int nt = nmax - 1 - n;
    const double *fxo = bufx[cur], *fyo = bufy[cur];
    double *fxn = bufx[1 - cur], *fyn = bufy[1 - cur];
    // d2_highorder_stencil + corner copies + fx/fy_calc_stencil_column (delnflux.py:128-213)
    fv3::launch3d(ctx, st, isc - nt, iec + 2 + nt, jsc - nt, jec + 2 + nt, 0, nk, FV_LAMBDA(int s, int i, int j, int k) {
      const bool hi = nord[k] > 0;
      const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
      if (!hi) {  // level keeps its first-order fluxes
        if (i >= isc && i <= iec + 1 && j >= jsc && j <= jec) fxn[o] = fxo[o];
        if (i >= isc && i <= iec && j >= jsc && j <= jec + 1) fyn[o] = fyo[o];
        return;
      }
      auto d2 = [&](int ii, int jj) {
        const int64_t p = O3(s, ii, jj, k);
        return (fxo[p] - fxo[p + 1] + fyo[p] - fyo[p + sj]) * m.rarea[O2(s, ii, jj)];
      };
      auto d2x = [&](int ii, int jj) {
        fv3::corner_x(g, s, ii, jj);
        return d2(ii, jj);
      };
      auto d2y = [&](int ii, int jj) {
        fv3::corner_y(g, s, ii, jj);
        return d2(ii, jj);
      };
      if (i <= iec + 1 + nt && j <= jec + nt) fxn[o] = -m.del6_v[o2] * (d2x(i - 1, j) - d2x(i, j));
      if (i <= iec + nt && j <= jec + 1 + nt) fyn[o] = -m.del6_u[o2] * (d2y(i, j - 1) - d2y(i, j));
    });
    cur = 1 - cur;
  }
  const double *fx2 = bufx[cur], *fy2 = bufy[cur];
  if (accumulate == 0) {
    fv3::launch3d(ctx, st, isc - nmax, iec + 2 + nmax, jsc - nmax, jec + 2 + nmax, 0, nk, FV_LAMBDA(int s, int i, int j, int k) {
      const int64_t o = O3(s, i, j, k);
      out_fx[o] = fx2[o];
      out_fy[o] = fy2[o];
    });
  } else {
    // add_diffusive_component / diffusive_damp (delnflux.py:215-238) on the (nx+1) x (ny+1) interface domain
    fv3::launch3d(ctx, st, isc, iec + 2, jsc, jec + 2, 0, nk, FV_LAMBDA(int s, int i, int j, int k) {
      const int64_t o = O3(s, i, j, k);
      if (accumulate == 1) {
        out_fx[o] = out_fx[o] + fx2[o];
        out_fy[o] = out_fy[o] + fy2[o];
      } else {
        out_fx[o] = out_fx[o] + 0.5 * damp[k] * (mass[o - 1] + mass[o]) * fx2[o];
        out_fy[o] = out_fy[o] + 0.5 * damp[k] * (mass[o - sj] + mass[o]) * fy2[o];
      }
    });
  }
}

}  // namespace

extern "C" {

int fv3_delnflux_nosg(fv3_ctx *ctx, const double *q, double *fx2, double *fy2, const double *damp_col,
                      const double *nord_col, int nmax, int nk, void *stream) {
  if (nmax > 2 || nmax < 0) {
    fv3::set_error("fv3_delnflux_nosg: nmax must be 0..2 (halo 3)");
    return -1;
  }
  delnflux_core(ctx, (cudaStream_t)stream, q, damp_col, nord_col, nmax, nk, false, fx2, fy2, 0, nullptr);
  return fv3::check_launch("fv3_delnflux_nosg");
}

int fv3_fvtp2d(fv3_ctx *ctx, const double *q, const double *crx, const double *cry, const double *xfx,
               const double *yfx, double *fx, double *fy, const double *x_mass_flux, const double *y_mass_flux,
               const double *mass, int hord, const double *nord_col, const double *damp_col, int nmax, int nk,
               void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  cudaStream_t st = (cudaStream_t)stream;
  const Idx x = make_idx(g);
  const int sj = g.sj;
  const int isc = x.isc, iec = x.iec, jsc = x.jsc, jec = x.jec, ied = x.ied, jed = x.jed;
  double *fy_in = fv3::scratch_field(ctx, S_FYIN), *fx_in = fv3::scratch_field(ctx, S_FXIN);
  double *q_i = fv3::scratch_field(ctx, S_QI), *q_j = fv3::scratch_field(ctx, S_QJ);
  const int ord_outer = hord, ord_inner = hord == 10 ? 8 : hord;
  const double *xu = x_mass_flux ? x_mass_flux : xfx, *yu = y_mass_flux ? y_mass_flux : yfx;

  // KA: inner sweeps on q with the cube corners remapped (fvtp2d.py:290-291,305-306)
  fv3::launch3d(ctx, st, 0, ied + 1, 0, jed + 1, 0, nk, FV_LAMBDA(int s, int i, int j, int k) {
    const int64_t o = O3(s, i, j, k);
    if (j >= jsc && j <= jec + 1) {
      const fv3::Edge1D e{fv3::on_south(g, s), fv3::on_north(g, s), jsc, jec};
      auto qy = [&](int jj) {
        int ii = i, j2 = jj;
        fv3::corner_y(g, s, ii, j2);
        return q[O3(s, ii, j2, k)];
      };
      auto dy = [&](int jj) { return m.dya[O2(s, i, jj)]; };
      fy_in[o] = fv3::ppm_flux(ord_inner, qy, dy, cry[o], j, e, true);
    }
    if (i >= isc && i <= iec + 1) {
      const fv3::Edge1D e{fv3::on_west(g, s), fv3::on_east(g, s), isc, iec};
      auto qx = [&](int ii) {
        int i2 = ii, jj = j;
        fv3::corner_x(g, s, i2, jj);
        return q[O3(s, i2, jj, k)];
      };
      auto dx = [&](int ii) { return m.dxa[O2(s, ii, j)]; };
      fx_in[o] = fv3::ppm_flux(ord_inner, qx, dx, crx[o], i, e, true);
    }
  });
  // KB: q advected along y / along x (q_i_stencil, q_j_stencil; fvtp2d.py:33-63)
  fv3::launch3d(ctx, st, 0, ied + 1, 0, jed + 1, 0, nk, FV_LAMBDA(int s, int i, int j, int k) {
    const int64_t o = O3(s, i, j, k);
    const double ar = m.area[O2(s, i, j)];
    if (j >= jsc && j <= jec) {
      const double f0 = yfx[o] * fy_in[o], f1 = yfx[o + sj] * fy_in[o + sj];
      q_i[o] = (q[o] * ar + f0 - f1) / (ar + yfx[o] - yfx[o + sj]);
    }
    if (i >= isc && i <= iec) {
      const double f0 = xfx[o] * fx_in[o], f1 = xfx[o + 1] * fx_in[o + 1];
      q_j[o] = (q[o] * ar + f0 - f1) / (ar + xfx[o] - xfx[o + 1]);
    }
  });
  // KC: outer sweeps and final fluxes (fvtp2d.py:66-93)
  fv3::launch3d(ctx, st, isc, iec + 2, jsc, jec + 2, 0, nk, FV_LAMBDA(int s, int i, int j, int k) {
    const int64_t o = O3(s, i, j, k);
    if (j <= jec) {
      const fv3::Edge1D e{fv3::on_west(g, s), fv3::on_east(g, s), isc, iec};
      auto qx = [&](int ii) { return q_i[O3(s, ii, j, k)]; };
      auto dx = [&](int ii) { return m.dxa[O2(s, ii, j)]; };
      const double outer = fv3::ppm_flux(ord_outer, qx, dx, crx[o], i, e, true);
      fx[o] = 0.5 * (outer + fx_in[o]) * xu[o];
    }
    if (i <= iec) {
      const fv3::Edge1D e{fv3::on_south(g, s), fv3::on_north(g, s), jsc, jec};
      auto qy = [&](int jj) { return q_j[O3(s, i, jj, k)]; };
      auto dy = [&](int jj) { return m.dya[O2(s, i, jj)]; };
      const double outer = fv3::ppm_flux(ord_outer, qy, dy, cry[o], j, e, true);
      fy[o] = 0.5 * (outer + fy_in[o]) * yu[o];
    }
  });
  if (damp_col != nullptr && nord_col != nullptr) {
    if (nmax > 2) {
      fv3::set_error("fv3_fvtp2d: nmax must be <= 2");
      return -1;
    }
    delnflux_core(ctx, st, q, damp_col, nord_col, nmax, nk, mass != nullptr, fx, fy, mass ? 2 : 1, mass);
  }
  return fv3::check_launch("fv3_fvtp2d");
}

}  // extern "C"
