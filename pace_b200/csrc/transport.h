// Plane-resident building blocks of the 2-D transport, shared by fvtp2d.cu (stand-alone transport, tracer sub-cycle)
// and d_sw.cu (the fused D-grid stages):
//   fvtp2d_plane   <- FiniteVolumeTransport.__call__ (fv3core/pace/fv3core/stencils/fvtp2d.py:235-346) incl.
//                     XPiecewiseParabolic / YPiecewiseParabolic (xppm.py:269-353, yppm.py)
//   delnflux_plane <- DelnFlux / DelnFluxNoSG (delnflux.py:59-238,1164-1261)
#pragma once
#include "common.h"
#include "plane.h"
#include "ppm.h"
#include "sweep.h"

namespace fv3 {


// ---- plane-resident transport (see plane.h) ------------------------------------------------------------------
// Shared-memory planes of the strip (Block::plane(n): resident rows x sj doubles, same (i, j) offsets as a global plane):
//   Q : q, cube corners filled for the y sweep, then for the x sweep; later q advected along y (q_i)
//   A : inner y-sweep interface values (fy_in); finally the y flux
//   B : inner x-sweep interface values (fx_in); finally the x flux
//   D : q advected along x (q_j)
//   T : per-sweep staging: PPM edge values al (hord 5/6) or limited slopes dm (hord 8), computed once per line
constexpr int FVTP_PLANES = 5;
struct PlaneArgs {
  const double *q, *crx, *cry, *xfx, *yfx, *xu, *yu;
  const double *add2d = nullptr;  // optional 2-D field (this subdomain's plane) added to q as it is loaded
};

// NC: the final-flux operands xu / yu are read-only for the whole kernel (read through the non-coherent path); false
// when the same kernel wrote them earlier (the fused d_sw stage keeps its mass fluxes in a scratch field).
// PRELOADED: the caller has already issued the bulk copy of q's resident rows into Q (it overlaps the caller's previous
// phase); only the wait for it happens here.
// LOOP: the two sweep passes are the two trips of one loop (one copy of each sweep in the kernel): measured faster for
// the hord-6 stages of d_sw (4004 -> 3915 us per call), slower for the hord-8 tracer sub-cycle (3412 -> 3753 us).
template <int MORD, bool NC = true, bool PRELOADED = false, bool LOOP = false>
FV_DEV void fvtp2d_plane(const fv3_geom &g, const fv3_grid &m, int s, int k, const Block &b, const PlaneArgs &a,
                        double *Q, double *A, double *B, double *D, double *T) {
  const int sj = g.sj, h = g.halo, nx = g.nx, ny = g.ny;
  const int isc = h, iec = h + nx - 1, jsc = h, jec = h + ny - 1, ied = iec + h, jed = jec + h;
  const int64_t ob = O3(s, 0, 0, k), o2b = O2(s, 0, 0);
  const double *q = a.q + ob, *crx = a.crx + ob, *cry = a.cry + ob, *xfx = a.xfx + ob, *yfx = a.yfx + ob;
  const double *dxa = m.dxa + o2b, *dya = m.dya + o2b, *area = m.area + o2b;
  const Edge1D ex{on_west(g, s), on_east(g, s), isc, iec};
  const Edge1D ey{on_south(g, s), on_north(g, s), jsc, jec};
  const int nwi = ied + 1, nwj = jed + 1;
  // this strip: fluxes on cell rows [ja, jb) and y faces [ja, jb]; q, fx_in and q_j on rows [rl, rh)
  const int ja = b.ja, jb = b.jb, rl = b.lo(0, h), rh = b.hi(nwj, h);
  b.prefetch_next_wave(a.q, g, k);
  {  // operands of the later phases: start their HBM -> L2 transfer now
    b.prefetch_rows(cry, sj);
    b.prefetch_rows(crx, sj);
    b.prefetch_rows(yfx, sj);
    b.prefetch_rows(xfx, sj);
    if (a.xu != a.xfx) b.prefetch_rows(a.xu + ob, sj);
    if (a.yu != a.yfx) b.prefetch_rows(a.yu + ob, sj);
  }
  // 1. load q: the strip's resident rows are one contiguous range of the global plane -> ONE bulk (TMA) copy; then the
  //    3x3 cube-corner halo blocks as copy_corners_y leaves them
  if (!PRELOADED) {
    b.bulk_begin(1, sj);
    b.bulk_rows(Q, q, sj);
  }
  b.bulk_wait();
  if (a.add2d) b.rect(0, nwi, rl, rh, [&](int i, int j) { Q[j * sj + i] = Q[j * sj + i] + a.add2d[j * sj + i]; });
  // the cube-corner halo blocks need their copy_corners_y / copy_corners_x forms only on tile-corner subdomains, and
  // only in the strips that hold rows of the y halo (block-uniform)
  const bool fix = (ex.lo || ex.hi) && (ey.lo || ey.hi) && (rl < jsc || rh > jec + 1);
  auto inner_y = make_sweep<MORD, false>(Q, sj, cry, dya, ey, 0, nwi, ja, jb, [&](int p, double val) { A[p] = val; });
  auto inner_x = make_sweep<MORD, true>(Q, sj, crx, dxa, ex, rl, rh - rl, isc, iec + 1, [&](int p, double val) { B[p] = val; });
  if (fix) {
    // The y sweeps of the x-halo columns want the corner blocks as copy_corners_y leaves them, the x sweeps of the
    // y-halo rows as copy_corners_x does.  Both forms at once: Q gets the x form, and the 2 x 3 halo columns are
    // copied into D (free until the transverse updates) with the y form of their corner cells — one small pass; the
    // y sweep then reads its halo columns from D.
    const int ncol = 2 * h, nrow = rh - rl;
    b.par(4 * h * h + ncol * nrow, [&](int t) {
      if (t < 4 * h * h) {
        const int c = t / (h * h), r = t - c * h * h, a1 = r / h, b1 = r - a1 * h;
        const int i = (c & 1) ? iec + 1 + a1 : a1, j = (c & 2) ? jec + 1 + b1 : b1;
        if (j < rl || j >= rh) return;
        int ii = i, jj = j;
        corner_x(g, s, ii, jj);
        if (ii != i || jj != j) Q[j * sj + i] = a.add2d ? q[jj * sj + ii] + a.add2d[jj * sj + ii] : q[jj * sj + ii];
      } else {
        const int t2 = t - 4 * h * h, jr = t2 / ncol, c = t2 - jr * ncol;
        const int i = c < h ? c : iec + 1 + (c - h), j = rl + jr;
        int ii = i, jj = j;
        corner_y(g, s, ii, jj);
        // a cell of a corner block comes from global memory (Q is being rewritten by the tasks above)
        if (j < jsc || j > jec)
          D[j * sj + i] = a.add2d ? q[jj * sj + ii] + a.add2d[jj * sj + ii] : q[jj * sj + ii];
        else
          D[j * sj + i] = Q[j * sj + i];
      }
    });
    inner_y.set_alt(D, isc, iec);
  }
  if (LOOP) {
  // 2 + 4 (inner sweeps), 5 (transverse updates), 6 + 7 (outer sweeps): the two sweep passes are the two trips of ONE loop,
  // so the kernel holds one copy of the x sweep and one of the y sweep (the interface value is stored as it is on the
  // first trip and turned into the flux on the second)
  const double *xu = a.xu + ob, *yu = a.yu + ob;
#pragma unroll 1
  for (int ph = 0; ph < 2; ++ph) {
    const bool outer = ph == 1;
    auto fin_y = [&](int p, double val) { A[p] = outer ? 0.5 * (val + A[p]) * (NC ? FV_LDG(yu + p) : yu[p]) : val; };
    auto fin_x = [&](int p, double val) { B[p] = outer ? 0.5 * (val + B[p]) * (NC ? FV_LDG(xu + p) : xu[p]) : val; };
    auto sy = make_sweep<MORD, false>(outer ? D : Q, sj, cry, dya, ey, outer ? isc : 0, outer ? nx : nwi, ja, jb, fin_y);
    auto sx = make_sweep<MORD, true>(Q, sj, crx, dxa, ex, outer ? ja : rl, outer ? jb - ja : rh - rl, isc, iec + 1, fin_x);
    if (!outer && fix) sy.set_alt(D, isc, iec);
    ppm_sweep_pair(b, sy, sx);
    if (!outer)
      b.rect(0, nwi, rl, rh, [&](int i, int j) {
        const int p = j * sj + i;
        const double qv = Q[p], ar = FV_LDG(area + p);
        if (j >= ja && j < jb) {
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
  }
    return;
  }
  // 2 + 4. both inner sweeps, one pass
  ppm_sweep_pair(b, inner_y, inner_x);
  // 5. transverse updates: q_i (into Q, owned rows) and q_j (into D, compute columns of the resident rows)
  b.rect(0, nwi, rl, rh, [&](int i, int j) {
    const int p = j * sj + i;
    const double qv = Q[p], ar = FV_LDG(area + p);
    if (j >= ja && j < jb) {
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
  // 6 + 7. outer x sweep on q_i (owned rows) -> x flux in place over fx_in; outer y sweep on q_j (compute columns) -> y
  // flux in place over fy_in: independent, one pass
  const double *xu = a.xu + ob, *yu = a.yu + ob;
  auto outer_x = make_sweep<MORD, true>(Q, sj, crx, dxa, ex, ja, jb - ja, isc, iec + 1,
                                        [&](int p, double val) { B[p] = 0.5 * (val + B[p]) * (NC ? FV_LDG(xu + p) : xu[p]); });
  auto outer_y = make_sweep<MORD, false>(D, sj, cry, dya, ey, isc, nx, ja, jb,
                                         [&](int p, double val) { A[p] = 0.5 * (val + A[p]) * (NC ? FV_LDG(yu + p) : yu[p]); });
  ppm_sweep_pair(b, outer_x, outer_y);
}

#ifndef FV3_HOSTSIM
__device__ __forceinline__ bool delnflux_plane_fast(const fv3_geom &g, const fv3_grid &m, int s, const Block &b,
                                                    const double *q, double dk, bool hi, int nmax, bool copy_q,
                                                    double *D2, double *FX, double *FY);
#endif
// ---- plane-resident del-n fluxes (DelnFlux / DelnFluxNoSG, delnflux.py:59-238,1164-1261) ------------------------
// D2: the field being differenced (damp * q, then the Laplacians of the previous fluxes), FX / FY: its fluxes.
// All nord iterations run in shared memory; on return FX / FY hold fx2 / fy2 on the strip's part of the interface
// domain: x fluxes on rows [ja, jb), y fluxes on faces [ja, jb].  q points at the global (s, k) plane.  An iteration
// whose results reach nt cells beyond the compute domain is evaluated nt rows beyond the strip.
FV_DEV void delnflux_plane(const fv3_geom &g, const fv3_grid &m, int s, const Block &b, const double *q, double dk,
                          bool hi, int nmax, bool copy_q, double *D2, double *FX, double *FY) {
#ifndef FV3_HOSTSIM
  if (delnflux_plane_fast(g, m, s, b, q, dk, hi, nmax, copy_q, D2, FX, FY)) return;
#endif
  const int sj = g.sj, h = g.halo, nx = g.nx;
  const int isc = h;
  const int ja = b.ja, jb = b.jb;
  const int64_t o2b = O2(s, 0, 0);
  const double *del6_u = m.del6_u + o2b, *del6_v = m.del6_v + o2b, *rarea = m.rarea + o2b;
  const int r = hi ? nmax : 0;
  // d2 = damp * q on cells [isc-r-1 .. iec+1+r] x [ja-r-1 .. jb+r]
  b.rect(isc - r - 1, isc + nx + r + 1, ja - r - 1, jb + r + 1, [&](int i, int j) {
    const int p = j * sj + i;
    const double v = q[p];
    D2[p] = copy_q ? v : dk * v;
  });
  // cube-corner remaps (copy_corners_x / _y) only matter where a difference touches a corner halo block: rows and
  // columns within one cell of the compute domain's ends; everywhere else the plain neighbours are read
  const int iec = isc + nx - 1, jsc = h, jec = h + g.ny - 1;
  const bool any_corner = (on_west(g, s) || on_east(g, s)) && (on_south(g, s) || on_north(g, s));
  auto near_corner = [&](int i, int j) { return any_corner && (i <= isc || i > iec) && (j <= jsc || j > jec); };
  auto d2x = [&](int ii, int jj) {
    corner_x(g, s, ii, jj);
    return D2[jj * sj + ii];
  };
  auto d2y = [&](int ii, int jj) {
    corner_y(g, s, ii, jj);
    return D2[jj * sj + ii];
  };
  b.rect(isc - r, isc + nx + r + 1, ja - r, jb + r + 1, [&](int i, int j) {
    const int p = j * sj + i;
    if (hi && near_corner(i, j)) {
      if (j < jb + r) FX[p] = del6_v[p] * (d2x(i - 1, j) - d2x(i, j));
      if (i < isc + nx + r) FY[p] = del6_u[p] * (d2y(i, j - 1) - d2y(i, j));
    } else {
      const double d0 = D2[p];
      if (j < jb + r) FX[p] = del6_v[p] * (D2[p - 1] - d0);
      if (i < isc + nx + r) FY[p] = del6_u[p] * (D2[p - sj] - d0);
    }
  });
  if (!hi) return;
  for (int n = 0; n < nmax; ++n) {
    const int nt = nmax - 1 - n;
    b.rect(isc - nt - 1, isc + nx + nt + 1, ja - nt - 1, jb + nt + 1, [&](int i, int j) {
      const int p = j * sj + i;
      D2[p] = (FX[p] - FX[p + 1] + FY[p] - FY[p + sj]) * rarea[p];
    });
    b.rect(isc - nt, isc + nx + nt + 1, ja - nt, jb + nt + 1, [&](int i, int j) {
      const int p = j * sj + i;
      if (near_corner(i, j)) {
        if (j < jb + nt) FX[p] = -del6_v[p] * (d2x(i - 1, j) - d2x(i, j));
        if (i < isc + nx + nt) FY[p] = -del6_u[p] * (d2y(i, j - 1) - d2y(i, j));
      } else {
        const double d0 = D2[p];
        if (j < jb + nt) FX[p] = -del6_v[p] * (D2[p - 1] - d0);
        if (i < isc + nx + nt) FY[p] = -del6_u[p] * (D2[p - sj] - d0);
      }
    });
  }
}

#ifndef FV3_HOSTSIM
// Same phases with a FIXED thread -> point assignment over the largest phase domain (3 cells / rows beyond the strip's
// compute part): the per-point flux coefficients (del6_v, del6_u) and the corner flags are fetched ONCE into
// registers, so the flux phases run on shared memory and registers only — no global-load wait at their start.  Same expressions as delnflux_plane; returns false when a thread would own more than DN_SLOTS
// points (the caller then uses the generic form).
constexpr int DN_SLOTS = 6;
__device__ __forceinline__ bool delnflux_plane_fast(const fv3_geom &g, const fv3_grid &m, int s, const Block &b,
                                                    const double *q, double dk, bool hi, int nmax, bool copy_q,
                                                    double *D2, double *FX, double *FY) {
  const int sj = g.sj, h = g.halo, nx = g.nx;
  const int isc = h, iec = isc + nx - 1, jsc = h, jec = h + g.ny - 1;
  const int ja = b.ja, jb = b.jb;
  const int w = nx + 6, npts = w * (jb - ja + 6), nthr = (int)blockDim.x;
  if (npts > DN_SLOTS * nthr) return false;
  const int64_t o2b = O2(s, 0, 0);
  const double *del6_u = m.del6_u + o2b, *del6_v = m.del6_v + o2b, *rarea = m.rarea + o2b;
  const bool any_corner = (on_west(g, s) || on_east(g, s)) && (on_south(g, s) || on_north(g, s));
  double dv[DN_SLOTS], du[DN_SLOTS];
  unsigned nc = 0;
  const float inv = 1.0f / (float)w;
  // slot n of this thread: point t = tid + n * nthr of the rectangle (i = -1000 for an empty slot: every domain test
  // fails); recomputed where needed instead of held in registers
  auto J = [&](int n) {
    const int t = (int)threadIdx.x + n * nthr;
    return (int)(((float)t + 0.5f) * inv);
  };
  auto I = [&](int n, int jr) {
    const int t = (int)threadIdx.x + n * nthr;
    return t < npts ? isc - 3 + (t - jr * w) : -1000;
  };
#pragma unroll
  for (int n = 0; n < DN_SLOTS; ++n) {
    const int jr = J(n), i = I(n, jr), j = ja - 3 + jr;
    dv[n] = du[n] = 0.0;
    if (i > -1000) {
      const int p = j * sj + i;
      dv[n] = FV_LDG(del6_v + p);
      du[n] = FV_LDG(del6_u + p);
      if (any_corner && (i <= isc || i > iec) && (j <= jsc || j > jec)) nc |= 1u << n;
    }
  }
  auto d2x = [&](int ii, int jj) {
    corner_x(g, s, ii, jj);
    return D2[jj * sj + ii];
  };
  auto d2y = [&](int ii, int jj) {
    corner_y(g, s, ii, jj);
    return D2[jj * sj + ii];
  };
  const int r = hi ? nmax : 0;
#pragma unroll
  for (int n = 0; n < DN_SLOTS; ++n) {
    const int jr = J(n), i = I(n, jr), j = ja - 3 + jr;
    if (i >= isc - r - 1 && i < isc + nx + r + 1 && j >= ja - r - 1 && j < jb + r + 1) {
      const int p = j * sj + i;
      const double v = q[p];
      D2[p] = copy_q ? v : dk * v;
    }
  }
  __syncthreads();
  // one flux phase: sign = +1 for the first differences, -1 for the iterations (written as in the generic form)
  auto flux = [&](int e, bool first) {
#pragma unroll
    for (int n = 0; n < DN_SLOTS; ++n) {
      const int jr = J(n), i = I(n, jr), j = ja - 3 + jr;
      if (i >= isc - e && i < isc + nx + e + 1 && j >= ja - e && j < jb + e + 1) {
        const int p = j * sj + i;
        const bool fxo = j < jb + e, fyo = i < isc + nx + e;
        if ((first ? hi : true) && ((nc >> n) & 1u)) {
          if (first) {
            if (fxo) FX[p] = dv[n] * (d2x(i - 1, j) - d2x(i, j));
            if (fyo) FY[p] = du[n] * (d2y(i, j - 1) - d2y(i, j));
          } else {
            if (fxo) FX[p] = -dv[n] * (d2x(i - 1, j) - d2x(i, j));
            if (fyo) FY[p] = -du[n] * (d2y(i, j - 1) - d2y(i, j));
          }
        } else {
          const double d0 = D2[p];
          if (first) {
            if (fxo) FX[p] = dv[n] * (D2[p - 1] - d0);
            if (fyo) FY[p] = du[n] * (D2[p - sj] - d0);
          } else {
            if (fxo) FX[p] = -dv[n] * (D2[p - 1] - d0);
            if (fyo) FY[p] = -du[n] * (D2[p - sj] - d0);
          }
        }
      }
    }
    __syncthreads();
  };
  flux(r, true);
  if (!hi) return true;
  for (int it = 0; it < nmax; ++it) {
    const int nt = nmax - 1 - it;
#pragma unroll
    for (int n = 0; n < DN_SLOTS; ++n) {
      const int jr = J(n), i = I(n, jr), j = ja - 3 + jr;
      if (i >= isc - nt - 1 && i < isc + nx + nt + 1 && j >= ja - nt - 1 && j < jb + nt + 1) {
        const int p = j * sj + i;
        D2[p] = (FX[p] - FX[p + 1] + FY[p] - FY[p + sj]) * FV_LDG(rarea + p);
      }
    }
    __syncthreads();
    flux(nt, false);
  }
  return true;
}
#endif


}  // namespace fv3
