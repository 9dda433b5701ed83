// 2-D flux-form finite-volume transport and the del-n diffusive fluxes.
//   fv3_fvtp2d        <- FiniteVolumeTransport.__call__ (fv3core/pace/fv3core/stencils/fvtp2d.py:235-346) incl.
//                        XPiecewiseParabolic / YPiecewiseParabolic (xppm.py:269-353, yppm.py) and, when a damping
//                        column is given, DelnFlux.__call__ (delnflux.py:1164-1207)
//   fv3_delnflux_nosg <- DelnFluxNoSG.__call__ (delnflux.py:1209-1261)
// 10 + (7 + 6*nmax) reference launches become 3 + (2 + nmax): the corner copies are index remaps at read time
// (ppm.h), the Laplacian d2 of each del-n iteration is recomputed from the previous fluxes instead of stored.
#include "common.h"
#include "plane.h"
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
    fv3::launch3d(ctx, st, isc - nmax, iec + 2 + nmax, jsc - nmax, jec + 2 + nmax, 0, nk, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
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
    fv3::launch3d(ctx, st, isc - nt, iec + 2 + nt, jsc - nt, jec + 2 + nt, 0, nk, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
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
    fv3::launch3d(ctx, st, isc - nmax, iec + 2 + nmax, jsc - nmax, jec + 2 + nmax, 0, nk, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
      const int64_t o = O3(s, i, j, k);
      out_fx[o] = fx2[o];
      out_fy[o] = fy2[o];
    });
  } else {
    // add_diffusive_component / diffusive_damp (delnflux.py:215-238) on the (nx+1) x (ny+1) interface domain
    fv3::launch3d(ctx, st, isc, iec + 2, jsc, jec + 2, 0, nk, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
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

// ---- plane-resident transport (see plane.h) ------------------------------------------------------------------
// Shared-memory planes (each PL = nj * sj doubles, same (i, j) offsets as a global plane):
//   Q : q, cube corners filled for the y sweep, then for the x sweep; later q advected along y (q_i)
//   A : inner y-sweep interface values (fy_in); finally the y flux
//   B : inner x-sweep interface values (fx_in); finally the x flux
//   D : q advected along x (q_j)
//   T : per-sweep staging: PPM edge values al (hord 5/6) or limited slopes dm (hord 8), computed once per line
constexpr int FVTP_PLANES = 5;
struct PlaneArgs {
  const double *q, *crx, *cry, *xfx, *yfx, *xu, *yu;
};

template <int MORD>
FV_HD void fvtp2d_plane(const fv3_geom &g, const fv3_grid &m, int s, int k, const fv3::Block &b, const PlaneArgs &a,
                        double *Q, double *A, double *B, double *D, double *T) {
  const int sj = g.sj, h = g.halo, nx = g.nx, ny = g.ny;
  const int isc = h, iec = h + nx - 1, jsc = h, jec = h + ny - 1, ied = iec + h, jed = jec + h;
  const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
  const double *q = a.q + ob, *crx = a.crx + ob, *cry = a.cry + ob, *xfx = a.xfx + ob, *yfx = a.yfx + ob;
  const double *dxa = m.dxa + o2b, *dya = m.dya + o2b, *area = m.area + o2b;
  const fv3::Edge1D ex{fv3::on_west(g, s), fv3::on_east(g, s), isc, iec};
  const fv3::Edge1D ey{fv3::on_south(g, s), fv3::on_north(g, s), jsc, jec};
  const int nwi = ied + 1, nwj = jed + 1;
  // staging ranges: al on faces start-1 .. end+2, dm on cells start-2 .. end+2
  const int st0 = MORD < 8 ? -1 : -2, stn = MORD < 8 ? 4 : 5;
  // 1. load q; 3x3 cube-corner halo blocks as copy_corners_y leaves them
  b.par2(nwi, nwj, [&](int i, int j) {
    int ii = i, jj = j;
    fv3::corner_y(g, s, ii, jj);
    Q[j * sj + i] = q[jj * sj + ii];
  });
  // 2. inner y sweep on q: all columns, faces jsc .. jec+1
  b.par2(nwi, ny + stn, [&](int i, int jr) {
    const int j = jsc + st0 + jr;
    auto qy = [&](int jj) { return Q[jj * sj + i]; };
    auto dy = [&](int jj) { return dya[jj * sj + i]; };
    T[j * sj + i] = fv3::ppm_stage<MORD>(qy, dy, j, ey);
  });
  b.par2(nwi, ny + 1, [&](int i, int jr) {
    const int j = jsc + jr, p = j * sj + i;
    auto qy = [&](int jj) { return Q[jj * sj + i]; };
    auto ty = [&](int jj) { return T[jj * sj + i]; };
    auto dy = [&](int jj) { return dya[jj * sj + i]; };
    A[p] = fv3::ppm_flux_staged<MORD>(qy, ty, dy, cry[p], j, ey);
  });
  // 3. cube-corner blocks as copy_corners_x leaves them
  b.par(4 * h * h, [&](int t) {
    const int c = t / (h * h), r = t - c * h * h, a1 = r / h, b1 = r - a1 * h;
    const int i = (c & 1) ? iec + 1 + a1 : a1, j = (c & 2) ? jec + 1 + b1 : b1;
    int ii = i, jj = j;
    fv3::corner_x(g, s, ii, jj);
    Q[j * sj + i] = q[jj * sj + ii];
  });
  // 4. inner x sweep on q: all rows, faces isc .. iec+1
  b.par2(nx + stn, nwj, [&](int ir, int j) {
    const int i = isc + st0 + ir;
    auto qx = [&](int ii) { return Q[j * sj + ii]; };
    auto dx = [&](int ii) { return dxa[j * sj + ii]; };
    T[j * sj + i] = fv3::ppm_stage<MORD>(qx, dx, i, ex);
  });
  b.par2(nx + 1, nwj, [&](int ir, int j) {
    const int i = isc + ir, p = j * sj + i;
    auto qx = [&](int ii) { return Q[j * sj + ii]; };
    auto tx = [&](int ii) { return T[j * sj + ii]; };
    auto dx = [&](int ii) { return dxa[j * sj + ii]; };
    B[p] = fv3::ppm_flux_staged<MORD>(qx, tx, dx, crx[p], i, ex);
  });
  // 5. transverse updates: q_i (into Q, compute rows) and q_j (into D, compute columns)
  b.par2(nwi, nwj, [&](int i, int j) {
    const int p = j * sj + i;
    const double qv = Q[p], ar = area[p];
    if (j >= jsc && j <= jec) {
      const double f0 = yfx[p] * A[p], f1 = yfx[p + sj] * A[p + sj];
      Q[p] = (qv * ar + f0 - f1) / (ar + yfx[p] - yfx[p + sj]);
    }
    if (i >= isc && i <= iec) {
      const double f0 = xfx[p] * B[p], f1 = xfx[p + 1] * B[p + 1];
      D[p] = (qv * ar + f0 - f1) / (ar + xfx[p] - xfx[p + 1]);
    }
  });
  // 6. outer x sweep on q_i (compute rows) -> x flux, in place over fx_in
  const double *xu = a.xu + ob, *yu = a.yu + ob;
  b.par2(nx + stn, ny, [&](int ir, int jr) {
    const int i = isc + st0 + ir, j = jsc + jr;
    auto qx = [&](int ii) { return Q[j * sj + ii]; };
    auto dx = [&](int ii) { return dxa[j * sj + ii]; };
    T[j * sj + i] = fv3::ppm_stage<MORD>(qx, dx, i, ex);
  });
  b.par2(nx + 1, ny, [&](int ir, int jr) {
    const int i = isc + ir, j = jsc + jr, p = j * sj + i;
    auto qx = [&](int ii) { return Q[j * sj + ii]; };
    auto tx = [&](int ii) { return T[j * sj + ii]; };
    auto dx = [&](int ii) { return dxa[j * sj + ii]; };
    const double outer = fv3::ppm_flux_staged<MORD>(qx, tx, dx, crx[p], i, ex);
    B[p] = 0.5 * (outer + B[p]) * xu[p];
  });
  // 7. outer y sweep on q_j (compute columns) -> y flux, in place over fy_in
  b.par2(nx, ny + stn, [&](int ir, int jr) {
    const int i = isc + ir, j = jsc + st0 + jr;
    auto qy = [&](int jj) { return D[jj * sj + i]; };
    auto dy = [&](int jj) { return dya[jj * sj + i]; };
    T[j * sj + i] = fv3::ppm_stage<MORD>(qy, dy, j, ey);
  });
  b.par2(nx, ny + 1, [&](int ir, int jr) {
    const int i = isc + ir, j = jsc + jr, p = j * sj + i;
    auto qy = [&](int jj) { return D[jj * sj + i]; };
    auto ty = [&](int jj) { return T[jj * sj + i]; };
    auto dy = [&](int jj) { return dya[jj * sj + i]; };
    const double outer = fv3::ppm_flux_staged<MORD>(qy, ty, dy, cry[p], j, ey);
    A[p] = 0.5 * (outer + A[p]) * yu[p];
  });
}

template <int MORD>
int fvtp2d_launch(const fv3_ctx *ctx, cudaStream_t st, PlaneArgs a, double *fx, double *fy, int nk) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int PL = g.nj * g.sj;
  return fv3::launch_planes(ctx, st, 0, nk, FVTP_PLANES * PL, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *Q = b.sm, *A = Q + PL, *B = A + PL, *D = B + PL, *T = D + PL;
    fvtp2d_plane<MORD>(g, m, s, k, b, a, Q, A, B, D, T);
    const int sj = g.sj, h = g.halo, nx1 = g.nx + 1;
    const int64_t ob = O3(s, 0, 0, k);
    b.par(nx1 * (g.ny + 1), [&](int t) {
      const int jr = t / nx1, i = h + (t - jr * nx1), j = h + jr, p = j * sj + i;
      if (jr < g.ny) fx[ob + p] = B[p];
      if (i - h < g.nx) fy[ob + p] = A[p];
    });
  });
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
  const double *xu = x_mass_flux ? x_mass_flux : xfx, *yu = y_mass_flux ? y_mass_flux : yfx;

  // inner sweeps, transverse updates, outer sweeps and final fluxes: ONE plane-resident kernel (fvtp2d.py:290-326)
  const PlaneArgs pa{q, crx, cry, xfx, yfx, xu, yu};
  const int mord = hord < 0 ? -hord : hord;
  int rc;
  if (mord == 8 || mord == 10)
    rc = fvtp2d_launch<8>(ctx, st, pa, fx, fy, nk);
  else if (mord == 5)
    rc = fvtp2d_launch<5>(ctx, st, pa, fx, fy, nk);
  else
    rc = fvtp2d_launch<6>(ctx, st, pa, fx, fy, nk);
  if (rc) return rc;
  if (damp_col != nullptr && nord_col != nullptr) {
    if (nmax > 2) {
      fv3::set_error("fv3_fvtp2d: nmax must be <= 2");
      return -1;
    }
    delnflux_core(ctx, st, q, damp_col, nord_col, nmax, nk, mass != nullptr, fx, fy, mass ? 2 : 1, mass);
  }
  return fv3::check_launch("fv3_fvtp2d");
}

// ---- TracerAdvection sub-cycle, all tracers in one launch (tracer_2d_1l.py:341-392) --------------------------
// Per plane: for every tracer, transport fluxes (hord_tr) stay in shared memory and are applied at once
// (apply_tracer_flux :138-156); dp2 (apply_mass_flux :115-135) is evaluated on the fly; after the last tracer
// dp2 is stored and, when another sub-cycle follows, swapped with dp1 (swap_dp :159-163).  Replaces
// 1 + 2*nq launches and 9 field passes per tracer by 1 launch and 2 passes per tracer (+ the shared flux fields,
// which stay L2-resident across the tracer loop).
int fv3_tracer_subcycle(fv3_ctx *ctx, double *const *tracers, int nq, double *dp1, double *dp2, const double *mfx,
                        const double *mfy, const double *cx, const double *cy, const double *xfx, const double *yfx,
                        int hord, int swap, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int mord = hord < 0 ? -hord : hord;
  if (mord != 8) {
    fv3::set_error("fv3_tracer_subcycle: only hord_tr = 8 is implemented");
    return -1;
  }
  const int PL = g.nj * g.sj;
  int rc = fv3::launch_planes(ctx, (cudaStream_t)stream, 0, g.nz, FVTP_PLANES * PL, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *Q = b.sm, *A = Q + PL, *B = A + PL, *D = B + PL, *T = D + PL;
    const int sj = g.sj, h = g.halo, nx = g.nx;
    const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
    const double *rarea = m.rarea + o2b;
    for (int n = 0; n < nq; ++n) {
      double *q = tracers[n];
      const PlaneArgs pa{q, cx, cy, xfx, yfx, mfx, mfy};
      fvtp2d_plane<8>(g, m, s, k, b, pa, Q, A, B, D, T);
      b.par(nx * g.ny, [&](int t) {
        const int jr = t / nx, p = (h + jr) * sj + h + (t - jr * nx);
        const int64_t o = ob + p;
        const double d1 = dp1[o];
        const double d2 = d1 + (mfx[o] - mfx[o + 1] + mfy[o] - mfy[o + sj]) * rarea[p];
        q[o] = (q[o] * d1 + (B[p] - B[p + 1] + A[p] - A[p + sj]) * rarea[p]) / d2;
      });
    }
    b.par(nx * g.ny, [&](int t) {
      const int jr = t / nx, p = (h + jr) * sj + h + (t - jr * nx);
      const int64_t o = ob + p;
      const double d1 = dp1[o];
      const double d2 = d1 + (mfx[o] - mfx[o + 1] + mfy[o] - mfy[o + sj]) * rarea[p];
      if (swap) {
        dp1[o] = d2;
        dp2[o] = d1;
      } else {
        dp2[o] = d2;
      }
    });
  });
  if (rc) return rc;
  return fv3::check_launch("fv3_tracer_subcycle");
}

}  // extern "C"
