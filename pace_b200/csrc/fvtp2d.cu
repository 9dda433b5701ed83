// 2-D flux-form finite-volume transport and the del-n diffusive fluxes.
//   fv3_fvtp2d        <- FiniteVolumeTransport.__call__ (fv3core/pace/fv3core/stencils/fvtp2d.py:235-346) incl.
//                        XPiecewiseParabolic / YPiecewiseParabolic (xppm.py:269-353, yppm.py) and, when a damping
//                        column is given, DelnFlux.__call__ (delnflux.py:1164-1207)
//   fv3_delnflux_nosg <- DelnFluxNoSG.__call__ (delnflux.py:1209-1261)
//   fv3_tracer_subcycle <- one sub-cycle of TracerAdvection.__call__ for all tracers (tracer_2d_1l.py:341-392)
// The 10 + (7 + 6*nmax) reference launches of a transport with damping are ONE strip-resident kernel (plane.h): the
// corner copies are index remaps at read time (ppm.h), every intermediate lives in shared memory.
#include "transport.h"


namespace {

using fv3::FVTP_PLANES;
using fv3::PlaneArgs;
using fv3::fvtp2d_plane;
using fv3::delnflux_plane;

// mode: 0 = transport fluxes only; 1 = fx += fx2 (DelnFlux without mass); 2 = fx += 0.5*damp*(mass[-1]+mass)*fx2
template <int MORD>
int fvtp2d_launch(const fv3_ctx *ctx, cudaStream_t st, PlaneArgs a, double *fx, double *fy, int nk, int mode,
                  const double *damp, const double *nord, int nmax, const double *mass) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  return fv3::launch_planes(ctx, st, 0, nk, FVTP_PLANES, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *Q = b.plane(0), *A = b.plane(1), *B = b.plane(2), *D = b.plane(3), *T = b.plane(4);
    const int64_t ob = O3(s, 0, 0, k);
    const int sj = g.sj, h = g.halo, nx = g.nx;
    if (mode == 2) b.prefetch_rows(mass + ob, sj);
    fvtp2d_plane<MORD>(g, m, s, k, b, a, Q, A, B, D, T);
    const int ja = b.ja, jb = b.jb, jt = b.jtop();
    if (mode == 0) {
      b.rect(h, h + nx + 1, ja, jt + 1, [&](int i, int j) {
        const int p = j * sj + i;
        if (j < jb) fx[ob + p] = B[p];
        if (i < h + nx) fy[ob + p] = A[p];
      });
      return;
    }
    // del-n damping fluxes of q, added to the transport fluxes (delnflux.py:1164-1207, 215-238)
    const double dk = damp[k];
    delnflux_plane(g, m, s, b, a.q + ob, dk, nord[k] > 0, nmax, mode == 2, Q, D, T);
    const double *ms = mass + ob;
    b.rect(h, h + nx + 1, ja, jt + 1, [&](int i, int j) {
      const int p = j * sj + i;
      if (mode == 1) {
        if (j < jb) fx[ob + p] = B[p] + D[p];
        if (i < h + nx) fy[ob + p] = A[p] + T[p];
      } else {
        if (j < jb) fx[ob + p] = B[p] + 0.5 * dk * (ms[p - 1] + ms[p]) * D[p];
        if (i < h + nx) fy[ob + p] = A[p] + 0.5 * dk * (ms[p - sj] + ms[p]) * T[p];
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
  int rc = fv3::launch_planes(ctx, (cudaStream_t)stream, 0, nk, 3, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *D2 = b.plane(0), *FX = b.plane(1), *FY = b.plane(2);
    const int64_t ob = O3(s, 0, 0, k);
    b.prefetch_next_wave(q, g, k);
    delnflux_plane(g, m, s, b, q + ob, damp_col[k], nord_col[k] > 0, nmax, false, D2, FX, FY);
    const int sj = g.sj, h = g.halo, nx = g.nx;
    b.rect(h, h + nx + 1, b.ja, b.jtop() + 1, [&](int i, int j) {
      const int p = j * sj + i;
      if (j < b.jb) fx2[ob + p] = FX[p];
      if (i < h + nx) fy2[ob + p] = FY[p];
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
  if (nq > 16) {
    fv3::set_error("fv3_tracer_subcycle: at most 16 tracers");
    return -1;
  }
  const fv3::StripGeom sg = fv3::strip_geometry(g, FVTP_PLANES, g.nz);
  double *side0 = fv3::scratch_field(ctx, 0);
  const int64_t side_stride = g.ss * g.n_sub;
  int rc = fv3::launch_planes(ctx, (cudaStream_t)stream, 0, g.nz, FVTP_PLANES, FV_LAMBDA(int s, int k, const fv3::Block &b) { FV_DEV_GM
    double *Q = b.plane(0), *A = b.plane(1), *B = b.plane(2), *D = b.plane(3), *T = b.plane(4);
    const int sj = g.sj, h = g.halo, nx = g.nx;
    const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
    const double *rarea = m.rarea + o2b;
    b.prefetch_next_wave(tracers[0], g, k);
    b.bulk_begin(1, sj);
    b.bulk_rows(Q, tracers[0] + ob, sj);
    for (int n = 0; n < nq; ++n) {
      double *q = tracers[n];
      if (n + 1 < nq) b.prefetch_rows(tracers[n + 1] + ob, sj);  // HBM -> L2 now, L2 -> shared memory after the sweeps
      const PlaneArgs pa{q, cx, cy, xfx, yfx, mfx, mfy};
      fvtp2d_plane<8, true, true>(g, m, s, k, b, pa, Q, A, B, D, T);
      // Q (q advected along y) is dead once the outer sweeps are done: the next tracer's rows arrive during the update
      if (n + 1 < nq) {
        b.bulk_begin(1, sj);
        b.bulk_rows(Q, tracers[n + 1] + ob, sj);
      }
      b.rect(h, h + nx, b.ja, b.jb, [&](int i, int j) {
        const int p = j * sj + i;
        const int64_t o = ob + p;
        const double d1 = dp1[o];
        const double d2 = d1 + (mfx[o] - mfx[o + 1] + mfy[o] - mfy[o + sj]) * rarea[p];
        // rows another strip of this plane reads as halo are parked in a side buffer (in-place update)
        const bool parked = (j < b.ja + h && b.ja > h) || (j >= b.jb - h && !b.last);
        (parked ? side0 + n * side_stride : q)[o] = (q[o] * d1 + (B[p] - B[p + 1] + A[p] - A[p + sj]) * rarea[p]) / d2;
      });
    }
    b.rect(h, h + nx, b.ja, b.jb, [&](int i, int j) {
      const int p = j * sj + i;
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
  if (sg.ns > 1) {
    // parked rows (3 either side of every interior strip boundary) back into the tracers
    const int h = g.halo, R = sg.rows_per_strip;
    fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, 0, 2 * h * (sg.ns - 1), 0, g.nz, FV_LAMBDA(int s, int i, int jj, int k) { FV_DEV_GM
      const int h2 = g.halo, bnd = jj / (2 * h2), j = h2 + (bnd + 1) * R - h2 + (jj - bnd * 2 * h2);
      if (j >= h2 + g.ny) return;  // a last strip shorter than the halo: rows beyond the compute domain were never parked
      const int64_t o = O3(s, i, j, k);
      for (int n = 0; n < nq; ++n) tracers[n][o] = side0[n * side_stride + o];
    });
  }
  return fv3::check_launch("fv3_tracer_subcycle");
}

}  // extern "C"
