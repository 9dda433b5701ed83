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
#include "sweep.h"

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
FV_DEV void fvtp2d_plane(const fv3_geom &g, const fv3_grid &m, int s, int k, const fv3::Block &b, const PlaneArgs &a,
                        double *Q, double *A, double *B, double *D, double *T) {
  const int sj = g.sj, h = g.halo, nx = g.nx, ny = g.ny;
  const int isc = h, iec = h + nx - 1, jsc = h, jec = h + ny - 1, ied = iec + h, jed = jec + h;
  const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
  const double *q = a.q + ob, *crx = a.crx + ob, *cry = a.cry + ob, *xfx = a.xfx + ob, *yfx = a.yfx + ob;
  const double *dxa = m.dxa + o2b, *dya = m.dya + o2b, *area = m.area + o2b;
  const fv3::Edge1D ex{fv3::on_west(g, s), fv3::on_east(g, s), isc, iec};
  const fv3::Edge1D ey{fv3::on_south(g, s), fv3::on_north(g, s), jsc, jec};
  const int nwi = ied + 1, nwj = jed + 1;
  {  // operands of the later phases: start their HBM -> L2 transfer now
    const int pl = g.nj * sj;
    b.prefetch_l2(cry, pl);
    b.prefetch_l2(crx, pl);
    b.prefetch_l2(yfx, pl);
    b.prefetch_l2(xfx, pl);
    if (a.xu != a.xfx) b.prefetch_l2(a.xu + ob, pl);
    if (a.yu != a.yfx) b.prefetch_l2(a.yu + ob, pl);
  }
  // 1. load q; 3x3 cube-corner halo blocks as copy_corners_y leaves them
  b.par2(nwi, nwj, [&](int i, int j) {
    int ii = i, jj = j;
    if ((i < isc || i > iec) && (j < jsc || j > jec)) fv3::corner_y(g, s, ii, jj);
    Q[j * sj + i] = FV_LDG(q + jj * sj + ii);
  });
  // 2. inner y sweep on q: all columns, faces jsc .. jec+1
  fv3::ppm_sweep<MORD, false>(b, Q, T, sj, cry, dya, ey, 0, nwi, [&](int p, double val) { A[p] = val; });
  // 3. cube-corner blocks as copy_corners_x leaves them
  b.par(4 * h * h, [&](int t) {
    const int c = t / (h * h), r = t - c * h * h, a1 = r / h, b1 = r - a1 * h;
    const int i = (c & 1) ? iec + 1 + a1 : a1, j = (c & 2) ? jec + 1 + b1 : b1;
    int ii = i, jj = j;
    fv3::corner_x(g, s, ii, jj);
    Q[j * sj + i] = q[jj * sj + ii];
  });
  // 4. inner x sweep on q: all rows, faces isc .. iec+1
  fv3::ppm_sweep<MORD, true>(b, Q, T, sj, crx, dxa, ex, 0, nwj, [&](int p, double val) { B[p] = val; });
  // 5. transverse updates: q_i (into Q, compute rows) and q_j (into D, compute columns)
  b.par2(nwi, nwj, [&](int i, int j) {
    const int p = j * sj + i;
    const double qv = Q[p], ar = FV_LDG(area + p);
    if (j >= jsc && j <= jec) {
      const double y0 = FV_LDG(yfx + p), y1 = FV_LDG(yfx + p + sj);
      const double f0 = y0 * A[p], f1 = y1 * A[p + sj];
      Q[p] = (qv * ar + f0 - f1) / (ar + y0 - y1);
    }
    if (i >= isc && i <= iec) {
      const double x0 = FV_LDG(xfx + p), x1 = FV_LDG(xfx + p + 1);
      const double f0 = x0 * B[p], f1 = x1 * B[p + 1];
      D[p] = (qv * ar + f0 - f1) / (ar + x0 - x1);
    }
  });
  // 6. outer x sweep on q_i (compute rows) -> x flux, in place over fx_in
  const double *xu = a.xu + ob, *yu = a.yu + ob;
  fv3::ppm_sweep<MORD, true>(b, Q, T, sj, crx, dxa, ex, jsc, ny,
                             [&](int p, double val) { B[p] = 0.5 * (val + B[p]) * FV_LDG(xu + p); });
  // 7. outer y sweep on q_j (compute columns) -> y flux, in place over fy_in
  fv3::ppm_sweep<MORD, false>(b, D, T, sj, cry, dya, ey, isc, nx,
                              [&](int p, double val) { A[p] = 0.5 * (val + A[p]) * FV_LDG(yu + p); });
}

// ---- plane-resident del-n fluxes (DelnFlux / DelnFluxNoSG, delnflux.py:59-238,1164-1261) ------------------------
// D2: the field being differenced (damp * q, then the Laplacians of the previous fluxes), FX / FY: its fluxes.
// All nord iterations run in shared memory; on return FX / FY hold fx2 / fy2 on the interface domain
// [isc..iec+1] x [jsc..jec(+1)].  q points at the global (s, k) plane.
FV_DEV void delnflux_plane(const fv3_geom &g, const fv3_grid &m, int s, const fv3::Block &b, const double *q, double dk,
                          bool hi, int nmax, bool copy_q, double *D2, double *FX, double *FY) {
  const int sj = g.sj, h = g.halo, nx = g.nx, ny = g.ny;
  const int isc = h, jsc = h;
  const int64_t o2b = O2(s, 0, 0);
  const double *del6_u = m.del6_u + o2b, *del6_v = m.del6_v + o2b, *rarea = m.rarea + o2b;
  const int r = hi ? nmax : 0;
  // d2 = damp * q on cells [isc-r-1 .. iec+1+r] x [jsc-r-1 .. jec+1+r]
  b.par2(nx + 2 * r + 2, ny + 2 * r + 2, [&](int ir, int jr) {
    const int p = (jsc - r - 1 + jr) * sj + isc - r - 1 + ir;
    const double v = q[p];
    D2[p] = copy_q ? v : dk * v;
  });
  auto d2x = [&](int ii, int jj) {
    if (hi) fv3::corner_x(g, s, ii, jj);
    return D2[jj * sj + ii];
  };
  auto d2y = [&](int ii, int jj) {
    if (hi) fv3::corner_y(g, s, ii, jj);
    return D2[jj * sj + ii];
  };
  b.par2(nx + 2 * r + 1, ny + 2 * r + 1, [&](int ir, int jr) {
    const int i = isc - r + ir, j = jsc - r + jr, p = j * sj + i;
    if (jr < ny + 2 * r) FX[p] = del6_v[p] * (d2x(i - 1, j) - d2x(i, j));
    if (ir < nx + 2 * r) FY[p] = del6_u[p] * (d2y(i, j - 1) - d2y(i, j));
  });
  if (!hi) return;
  for (int n = 0; n < nmax; ++n) {
    const int nt = nmax - 1 - n;
    b.par2(nx + 2 * nt + 2, ny + 2 * nt + 2, [&](int ir, int jr) {
      const int p = (jsc - nt - 1 + jr) * sj + isc - nt - 1 + ir;
      D2[p] = (FX[p] - FX[p + 1] + FY[p] - FY[p + sj]) * rarea[p];
    });
    b.par2(nx + 2 * nt + 1, ny + 2 * nt + 1, [&](int ir, int jr) {
      const int i = isc - nt + ir, j = jsc - nt + jr, p = j * sj + i;
      int ia = i - 1, ja = j, ib = i, jb = j;
      fv3::corner_x(g, s, ia, ja);
      fv3::corner_x(g, s, ib, jb);
      if (jr < ny + 2 * nt) FX[p] = -del6_v[p] * (D2[ja * sj + ia] - D2[jb * sj + ib]);
      ia = i, ja = j - 1, ib = i, jb = j;
      fv3::corner_y(g, s, ia, ja);
      fv3::corner_y(g, s, ib, jb);
      if (ir < nx + 2 * nt) FY[p] = -del6_u[p] * (D2[ja * sj + ia] - D2[jb * sj + ib]);
    });
  }
}

// mode: 0 = transport fluxes only; 1 = fx += fx2 (DelnFlux without mass); 2 = fx += 0.5*damp*(mass[-1]+mass)*fx2
template <int MORD>
int fvtp2d_launch(const fv3_ctx *ctx, cudaStream_t st, PlaneArgs a, double *fx, double *fy, int nk, int mode,
                  const double *damp, const double *nord, int nmax, const double *mass) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int PL = g.nj * g.sj;
  return fv3::launch_planes(ctx, st, 0, nk, FVTP_PLANES * PL, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *Q = b.sm, *A = Q + PL, *B = A + PL, *D = B + PL, *T = D + PL;
    const int64_t ob = O3(s, 0, 0, k);
    if (mode == 2) b.prefetch_l2(mass + ob, PL);
    fvtp2d_plane<MORD>(g, m, s, k, b, a, Q, A, B, D, T);
    const int sj = g.sj, h = g.halo, nx = g.nx, ny = g.ny;
    if (mode == 0) {
      b.par2(nx + 1, ny + 1, [&](int ir, int jr) {
        const int p = (h + jr) * sj + h + ir;
        if (jr < ny) fx[ob + p] = B[p];
        if (ir < nx) fy[ob + p] = A[p];
      });
      return;
    }
    // del-n damping fluxes of q, added to the transport fluxes (delnflux.py:1164-1207, 215-238)
    const double dk = damp[k];
    delnflux_plane(g, m, s, b, a.q + ob, dk, nord[k] > 0, nmax, mode == 2, Q, D, T);
    const double *ms = mass + ob;
    b.par2(nx + 1, ny + 1, [&](int ir, int jr) {
      const int p = (h + jr) * sj + h + ir;
      if (mode == 1) {
        if (jr < ny) fx[ob + p] = B[p] + D[p];
        if (ir < nx) fy[ob + p] = A[p] + T[p];
      } else {
        if (jr < ny) fx[ob + p] = B[p] + 0.5 * dk * (ms[p - 1] + ms[p]) * D[p];
        if (ir < nx) fy[ob + p] = A[p] + 0.5 * dk * (ms[p - sj] + ms[p]) * T[p];
      }
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
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int PL = g.nj * g.sj;
  int rc = fv3::launch_planes(ctx, (cudaStream_t)stream, 0, nk, 3 * PL, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *D2 = b.sm, *FX = D2 + PL, *FY = FX + PL;
    const int64_t ob = O3(s, 0, 0, k);
    delnflux_plane(g, m, s, b, q + ob, damp_col[k], nord_col[k] > 0, nmax, false, D2, FX, FY);
    const int sj = g.sj, h = g.halo, nx = g.nx, ny = g.ny;
    b.par2(nx + 1, ny + 1, [&](int ir, int jr) {
      const int p = (h + jr) * sj + h + ir;
      if (jr < ny) fx2[ob + p] = FX[p];
      if (ir < nx) fy2[ob + p] = FY[p];
    });
  });
  if (rc) return rc;
  return fv3::check_launch("fv3_delnflux_nosg");
}

int fv3_fvtp2d(fv3_ctx *ctx, const double *q, const double *crx, const double *cry, const double *xfx,
               const double *yfx, double *fx, double *fy, const double *x_mass_flux, const double *y_mass_flux,
               const double *mass, int hord, const double *nord_col, const double *damp_col, int nmax, int nk,
               void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const double *xu = x_mass_flux ? x_mass_flux : xfx, *yu = y_mass_flux ? y_mass_flux : yfx;

  // inner sweeps, transverse updates, outer sweeps, final fluxes and (optionally) the del-n damping fluxes:
  // ONE plane-resident kernel (fvtp2d.py:290-346)
  const PlaneArgs pa{q, crx, cry, xfx, yfx, xu, yu};
  const int mord = hord < 0 ? -hord : hord;
  int mode = 0;
  if (damp_col != nullptr && nord_col != nullptr) {
    if (nmax > 2) {
      fv3::set_error("fv3_fvtp2d: nmax must be <= 2");
      return -1;
    }
    mode = mass ? 2 : 1;
  }
  const double *ms = mass ? mass : q;
  int rc;
  if (mord == 8 || mord == 10)
    rc = fvtp2d_launch<8>(ctx, st, pa, fx, fy, nk, mode, damp_col, nord_col, nmax, ms);
  else if (mord == 5)
    rc = fvtp2d_launch<5>(ctx, st, pa, fx, fy, nk, mode, damp_col, nord_col, nmax, ms);
  else
    rc = fvtp2d_launch<6>(ctx, st, pa, fx, fy, nk, mode, damp_col, nord_col, nmax, ms);
  if (rc) return rc;
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
    b.prefetch_l2(dp1 + ob, PL);
    for (int n = 0; n < nq; ++n) {
      double *q = tracers[n];
      if (n + 1 < nq) b.prefetch_l2(tracers[n + 1] + ob, PL);
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
