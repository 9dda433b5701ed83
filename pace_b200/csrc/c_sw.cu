// C-grid shallow-water half step.
//   fv3_c_sw  <-  CGridShallowWaterDynamics.__call__ (fv3core/pace/fv3core/stencils/c_sw.py:607-766) including
//                 DGrid2AGrid2CGridVectors.__call__ (d2a2c_vect.py:547-655) and the corner fills it uses
//                 (stencils/pace/stencils/corners.py:130-304).
// The reference's 24 stencil launches are regrouped into 6 launches; region-restricted statements become
// per-subdomain tile-edge predicates (geom.edge).  Values, not statement order, are reproduced: every point is
// computed with the formula that "wins" in the reference's statement sequence.
#include "common.h"

namespace {

constexpr double A1 = 9.0 / 16.0;
constexpr double A2 = -1.0 / 16.0;
constexpr double C1 = -2.0 / 14.0;
constexpr double C2 = 11.0 / 14.0;
constexpr double C3 = 5.0 / 14.0;
constexpr double BIG = 1e30;

FV_HD double contravariant(double v1, double v2, double cosa, double rsin2) { return (v1 - v2 * cosa) * rsin2; }

// fill_corners_2cells_x / _y (corners.py:130-166,235-270) as read-time remaps: the cell (i, j) of a cube-corner halo block
// reads the cell the in-place fill would have copied into it
FV_HD void fill2_x(const fv3_geom &g, int s, int &i, int &j) {
  const int isc = g.halo, iec = g.halo + g.nx - 1, jsc = g.halo, jec = g.halo + g.ny - 1;
  const bool south = j == jsc - 1, north = j == jec + 1;
  if (!(south || north)) return;
  if (!(south ? fv3::on_south(g, s) : fv3::on_north(g, s))) return;
  const int ys = south ? -1 : 1;
  if (fv3::on_west(g, s) && (i == isc - 1 || i == isc - 2)) {
    j = j - ys * (isc - i);
    i = isc - 1;
  } else if (fv3::on_east(g, s) && (i == iec + 1 || i == iec + 2)) {
    j = j - ys * (i - iec);
    i = iec + 1;
  }
}
FV_HD void fill2_y(const fv3_geom &g, int s, int &i, int &j) {
  const int isc = g.halo, iec = g.halo + g.nx - 1, jsc = g.halo, jec = g.halo + g.ny - 1;
  const bool west = i == isc - 1, east = i == iec + 1;
  if (!(west || east)) return;
  if (!(west ? fv3::on_west(g, s) : fv3::on_east(g, s))) return;
  const int xs = west ? -1 : 1;
  if (fv3::on_south(g, s) && (j == jsc - 1 || j == jsc - 2)) {
    i = i - xs * (jsc - j);
    j = jsc - 1;
  } else if (fv3::on_north(g, s) && (j == jec + 1 || j == jec + 2)) {
    i = i - xs * (j - jec);
    j = jec + 1;
  }
}

}  // namespace

extern "C" {

int fv3_c_sw(fv3_ctx *ctx, double *delp, double *pt, const double *u, const double *v, double *w, double *uc,
             double *vc, double *ua, double *va, double *ut, double *vt, double *divgd, double *omga, double *delpc,
             double *ptc, double dt2, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  cudaStream_t st = (cudaStream_t)stream;
  const int h = g.halo, nz = g.nz;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  const int ied = iec + h, jed = jec + h;
  const int sj = g.sj;
  double *utmp = fv3::scratch_field(ctx, 0), *vtmp = fv3::scratch_field(ctx, 1);
  double *ke = fv3::scratch_field(ctx, 5);
  double *uc0 = fv3::scratch_field(ctx, 2), *vc0 = fv3::scratch_field(ctx, 3);  // d2a2c winds before the update
  int npt = 4;
  if (npt > g.nx - 1 || npt > g.ny - 1) npt = 0;
  const int off = npt == 0 ? -1 : 3;
  const int nord = ctx->c.nord;

  // K1: utmp/vtmp (d2a2c_vect.py:19-65) on the full domain; zero delpc/ptc (c_sw.py:19-27) outside the transport's domain
  fv3::launch3d(ctx, st, 0, ied + 1, 0, jed + 1, 0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int64_t o = O3(s, i, j, k);
    if (i < isc - 1 || i > iec + 1 || j < jsc - 1 || j > jec + 1) {  // elsewhere the transport launch stores them
      delpc[o] = 0.0;
      ptc[o] = 0.0;
    }
    const bool avg = (S && j < jsc + off) || (N && j > jec - off) || (W && i < isc + off) || (E && i > iec - off);
    double ut_ = BIG, vt_ = BIG;
    if (avg) {
      ut_ = 0.5 * (u[o] + u[o + sj]);
      vt_ = 0.5 * (v[o] + v[o + 1]);
    } else {
      const int js1 = S ? npt + 2 : jsc - 1, je1 = N ? jec + 1 - npt : jec + 1;
      const int is1 = W ? npt + 2 : 0, ie1 = E ? iec + 1 - npt : ied;
      const int is2 = W ? npt + 2 : isc - 1, ie2 = E ? iec + 1 - npt : iec + 1;
      const int js2 = S ? npt + 2 : 0, je2 = N ? jec + 1 - npt : jed;
      if (i >= is1 && i <= ie1 && j >= js1 && j <= je1)
        ut_ = A2 * (u[o - sj] + u[o + 2 * sj]) + A1 * (u[o] + u[o + sj]);
      if (i >= is2 && i <= ie2 && j >= js2 && j <= je2) vt_ = A2 * (v[o - 1] + v[o + 2]) + A1 * (v[o] + v[o + 1]);
    }
    utmp[o] = ut_;
    vtmp[o] = vt_;
    // contravariant A-grid winds on compute + 2 (d2a2c_vect.py:68-78), from the values just formed
    if (i >= isc - 2 && i <= iec + 2 && j >= jsc - 2 && j <= jec + 2) {
      const int64_t o2 = O2(s, i, j);
      const double cs = m.cosa_s[o2], r2 = m.rsin2[o2];
      ua[o] = contravariant(ut_, vt_, cs, r2);
      va[o] = contravariant(vt_, ut_, cs, r2);
    }
  });

  // K2b: corner fills of utmp (3 cells), ua (2 cells) in x and vtmp, va in y (d2a2c_vect.py:81-88,157-164);
  // read and written point sets are disjoint, so x and y fills share one launch.
  fv3::launch3d(ctx, st, 0, 12, 0, 2, 0, nz, FV_LAMBDA(int s, int id, int dir, int k) { FV_DEV_GM
    const int corner = id / 3, d = id % 3 + 1;  // corner: 0 sw, 1 se, 2 nw, 3 ne
    const bool west = (corner == 0 || corner == 2), south = (corner < 2);
    if (!((west ? fv3::on_west(g, s) : fv3::on_east(g, s)) && (south ? fv3::on_south(g, s) : fv3::on_north(g, s))))
      return;
    const double mult = (corner == 0 || corner == 3) ? -1.0 : 1.0;
    const int ic = west ? isc - 1 : iec + 1;   // first halo column
    const int jc = south ? jsc - 1 : jec + 1;  // first halo row
    const int xs = west ? -1 : 1, ys = south ? -1 : 1;
    if (dir == 0) {
      // q(ic + xs*(d-1), jc) = mult * qc(ic, jc - ys*d)
      const int64_t od = O3(s, ic + xs * (d - 1), jc, k), os = O3(s, ic, jc - ys * d, k);
      utmp[od] = mult * vtmp[os];
      if (d <= 2) ua[od] = mult * va[os];
    } else {
      // q(ic, jc + ys*(d-1)) = mult * qc(ic - xs*d, jc)
      const int64_t od = O3(s, ic, jc + ys * (d - 1), k), os = O3(s, ic - xs * d, jc, k);
      vtmp[od] = mult * utmp[os];
      if (d <= 2) va[od] = mult * ua[os];
    }
  });

  // K3 + K4: C-grid winds (into uc0, vc0: the wind update below reads them while it stores uc, vc) and geo-adjusted
  // contravariant fluxes ut, vt (d2a2c_vect.py:91-228, c_sw.py:156-199); divergence at cell corners
  fv3::launch3d(ctx, st, isc - 1, iec + 3, jsc - 1, jec + 3, 0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    if (j <= jec + 1) {
      double ucv, utv;
      int e = 0;  // 1: cubic, 2: edge interpolation, 3: cubic reversed
      if (W && i >= isc - 1 && i <= isc + 1) e = i - (isc - 1) + 1;
      if (E && i >= iec && i <= iec + 2) e = i - iec + 1;
      if (e == 2) {
        const double t1 = m.dxa[o2 - 2] + m.dxa[o2 - 1], t2 = m.dxa[o2] + m.dxa[o2 + 1];
        const double n1 = (t1 + m.dxa[o2 - 1]) * ua[o - 1] - m.dxa[o2 - 1] * ua[o - 2];
        const double n2 = (t1 + m.dxa[o2]) * ua[o] - m.dxa[o2] * ua[o + 1];
        utv = 0.5 * (n1 / t1 + n2 / t2);
        ucv = utv > 0 ? utv * m.sin_sg3[o2 - 1] : utv * m.sin_sg1[o2];
      } else {
        if (e == 1)
          ucv = C1 * utmp[o - 2] + C2 * utmp[o - 1] + C3 * utmp[o];
        else if (e == 3)
          ucv = C1 * utmp[o + 1] + C2 * utmp[o] + C3 * utmp[o - 1];
        else
          ucv = A2 * (utmp[o - 2] + utmp[o + 1]) + A1 * (utmp[o - 1] + utmp[o]);
        utv = contravariant(ucv, v[o], m.cosa_u[o2], m.rsin_u[o2]);
      }
      uc0[o] = ucv;
      ut[o] = utv > 0 ? dt2 * utv * m.dy[o2] * m.sin_sg3[o2 - 1] : dt2 * utv * m.dy[o2] * m.sin_sg1[o2];
    }
    if (i <= iec + 1) {
      double vcv, vtv;
      int e = 0;
      if (S && j >= jsc - 1 && j <= jsc + 1) e = j - (jsc - 1) + 1;
      if (N && j >= jec && j <= jec + 2) e = j - jec + 1;
      if (e == 2) {
        const double t1 = m.dya[o2 - 2 * sj] + m.dya[o2 - sj], t2 = m.dya[o2] + m.dya[o2 + sj];
        const double n1 = (t1 + m.dya[o2 - sj]) * va[o - sj] - m.dya[o2 - sj] * va[o - 2 * sj];
        const double n2 = (t1 + m.dya[o2]) * va[o] - m.dya[o2] * va[o + sj];
        vtv = 0.5 * (n1 / t1 + n2 / t2);
        vcv = vtv > 0 ? vtv * m.sin_sg4[o2 - sj] : vtv * m.sin_sg2[o2];
      } else {
        if (e == 1)
          vcv = C1 * vtmp[o - 2 * sj] + C2 * vtmp[o - sj] + C3 * vtmp[o];
        else if (e == 3)
          vcv = C1 * vtmp[o + sj] + C2 * vtmp[o] + C3 * vtmp[o - sj];
        else
          vcv = A2 * (vtmp[o - 2 * sj] + vtmp[o + sj]) + A1 * (vtmp[o - sj] + vtmp[o]);
        vtv = contravariant(vcv, u[o], m.cosa_v[o2], m.rsin_v[o2]);
      }
      vc0[o] = vcv;
      vt[o] = vtv > 0 ? dt2 * vtv * m.dx[o2] * m.sin_sg4[o2 - sj] : dt2 * vtv * m.dx[o2] * m.sin_sg2[o2];
    }
    // divergence at cell corners (c_sw.py:31-154): reads u, v, ua, va like the winds above, nothing the launch writes
    if (nord > 0 && i >= isc && i <= iec + 1 && j >= jsc && j <= jec + 1) {
      auto uf = [&](int ii, int jj) {
        const int64_t o = O3(s, ii, jj, k), o2 = O2(s, ii, jj);
        if ((S && jj == jsc) || (N && jj == jec + 1)) return u[o] * m.dyc[o2] * 0.5 * (m.sin_sg4[o2 - sj] + m.sin_sg2[o2]);
        return (u[o] - 0.25 * (va[o - sj] + va[o]) * (m.cos_sg4[o2 - sj] + m.cos_sg2[o2])) * m.dyc[o2] * 0.5 *
               (m.sin_sg4[o2 - sj] + m.sin_sg2[o2]);
      };
      auto vf = [&](int ii, int jj) {
        const int64_t o = O3(s, ii, jj, k), o2 = O2(s, ii, jj);
        if ((W && ii == isc) || (E && ii == iec + 1)) return v[o] * m.dxc[o2] * 0.5 * (m.sin_sg3[o2 - 1] + m.sin_sg1[o2]);
        return (v[o] - 0.25 * (ua[o - 1] + ua[o]) * (m.cos_sg3[o2 - 1] + m.cos_sg1[o2])) * m.dxc[o2] * 0.5 *
               (m.sin_sg3[o2 - 1] + m.sin_sg1[o2]);
      };
      const bool ci = (W && i == isc) || (E && i == iec + 1);
      const double rc = m.rarea_c[O2(s, i, j)];
      double d;
      if (ci && S && j == jsc)
        d = (-vf(i, j) + uf(i - 1, j) - uf(i, j)) * rc;
      else if (ci && N && j == jec + 1)
        d = (vf(i, j - 1) + uf(i - 1, j) - uf(i, j)) * rc;
      else
        d = (vf(i, j - 1) - vf(i, j) + uf(i - 1, j) - uf(i, j)) * rc;
      divgd[O3(s, i, j, k)] = d;
    }
  });

  // K5 + K6: first-order upwind transport of delp, pt, w (c_sw.py:229-345), upstream kinetic energy and vorticity
  // (:347-364).  The x fluxes of a cell's two faces are formed on the fly (two multiplies each) instead of going through
  // three scratch fields; the in-place 2-cell corner fills of delp, pt, w (corners.py:130-166,235-270) are read-time
  // index remaps — the x fluxes see the cells as fill_corners_2cells_x leaves them, everything else as
  // fill_corners_2cells_y does — and the cells the reference leaves modified are written by a small launch afterwards.
  fv3::launch3d(ctx, st, isc - 1, iec + 2, jsc - 1, jec + 2, 0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    const bool near_corner = (W || E) && (S || N) && (i <= isc || i >= iec) && (j <= jsc || j >= jec);
    auto rd = [&](const double *q, int ii, int jj, bool xfill) {
      if (near_corner) {
        if (xfill)
          fill2_x(g, s, ii, jj);
        else
          fill2_y(g, s, ii, jj);
      }
      return q[O3(s, ii, jj, k)];
    };
    double fx1a, fxa, fx2a, fx1b, fxb, fx2b;
    {
      const double utc = ut[o];
      const int iu = utc > 0.0 ? i - 1 : i;
      fx1a = utc * rd(delp, iu, j, true);
      fxa = fx1a * rd(pt, iu, j, true);
      fx2a = fx1a * rd(w, iu, j, true);
    }
    {
      const double utc = ut[o + 1];
      const int iu = utc > 0.0 ? i : i + 1;
      fx1b = utc * rd(delp, iu, j, true);
      fxb = fx1b * rd(pt, iu, j, true);
      fx2b = fx1b * rd(w, iu, j, true);
    }
    double fy1a, fya, fy2a, fy1b, fyb, fy2b;
    {
      const double vtc = vt[o];
      const int ju = vtc > 0.0 ? j - 1 : j;
      fy1a = vtc * rd(delp, i, ju, false);
      fya = fy1a * rd(pt, i, ju, false);
      fy2a = fy1a * rd(w, i, ju, false);
    }
    {
      const double vtc = vt[o + sj];
      const int ju = vtc > 0.0 ? j : j + 1;
      fy1b = vtc * rd(delp, i, ju, false);
      fyb = fy1b * rd(pt, i, ju, false);
      fy2b = fy1b * rd(w, i, ju, false);
    }
    const double ra = m.rarea[o2];
    const double dp0 = rd(delp, i, j, false), pt0 = rd(pt, i, j, false), w0 = rd(w, i, j, false);
    const double dpc = dp0 + (fx1a - fx1b + fy1a - fy1b) * ra;
    delpc[o] = dpc;
    ptc[o] = (pt0 * dp0 + (fxa - fxb + fya - fyb) * ra) / dpc;
    omga[o] = (w0 * dp0 + (fx2a - fx2b + fy2a - fy2b) * ra) / dpc;
    const double uav = ua[o], vav = va[o];
    double kev = uav > 0.0 ? uc0[o] : uc0[o + 1];
    double vo = vav > 0.0 ? vc0[o] : vc0[o + sj];
    if ((S && j == jsc - 1) || (N && j == jec)) vo = vav <= 0.0 ? vo * m.sin_sg4[o2] + u[o + sj] * m.cos_sg4[o2] : vo;
    if ((S && j == jsc) || (N && j == jec + 1)) vo = vav > 0.0 ? vo * m.sin_sg2[o2] + u[o] * m.cos_sg2[o2] : vo;
    if ((E && i == iec) || (W && i == isc - 1)) kev = uav <= 0.0 ? kev * m.sin_sg3[o2] + v[o + 1] * m.cos_sg3[o2] : kev;
    if ((E && i == iec + 1) || (W && i == isc)) kev = uav > 0.0 ? kev * m.sin_sg1[o2] + v[o] * m.cos_sg1[o2] : kev;
    ke[o] = 0.5 * dt2 * (uav * kev + vav * vo);
  });
  // what the in-place corner fills leave behind: per tile corner (ic, jc) and (ic, jc + ys) hold their y fill,
  // (ic + xs, jc) its x fill (the sources are cells no fill touches)
  fv3::launch3d(ctx, st, 0, 12, 0, 3, 0, nz, FV_LAMBDA(int s, int id, int f, int k) { FV_DEV_GM
    const int corner = id / 3, c = id % 3;
    const bool west = (corner == 0 || corner == 2), south = (corner < 2);
    if (!((west ? fv3::on_west(g, s) : fv3::on_east(g, s)) && (south ? fv3::on_south(g, s) : fv3::on_north(g, s))))
      return;
    double *q = f == 0 ? delp : (f == 1 ? pt : w);
    const int ic = west ? isc - 1 : iec + 1, jc = south ? jsc - 1 : jec + 1;
    const int xs = west ? -1 : 1, ys = south ? -1 : 1;
    if (c == 0)
      q[O3(s, ic, jc, k)] = q[O3(s, ic - xs, jc, k)];
    else if (c == 1)
      q[O3(s, ic, jc + ys, k)] = q[O3(s, ic - 2 * xs, jc, k)];
    else
      q[O3(s, ic + xs, jc, k)] = q[O3(s, ic, jc - 2 * ys, k)];
  });

  // K7 + K8: absolute vorticity at cell corners (c_sw.py:367-408) and the C-grid wind update (:411-480) in one launch.
  // A wind needs the vorticity of ONE of two corners (the upwind one): it is evaluated on the fly from uc0 / vc0 instead
  // of going through a field; the launch covers the whole domain the d2a2c winds were formed on and copies them where
  // no update applies.
  fv3::launch3d(ctx, st, isc - 1, iec + 3, jsc - 1, jec + 3, 0, nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const bool W = fv3::on_west(g, s), E = fv3::on_east(g, s), S = fv3::on_south(g, s), N = fv3::on_north(g, s);
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    auto vort = [&](int ii, int jj) {
      const int64_t oo = O3(s, ii, jj, k), oo2 = O2(s, ii, jj);
      const double fxv = m.dxc[oo2] * uc0[oo], fyv = m.dyc[oo2] * vc0[oo];
      const double fx1v = m.dxc[oo2 - sj] * uc0[oo - sj], fy1v = m.dyc[oo2 - 1] * vc0[oo - 1];
      double vc_ = fx1v - fxv - fy1v + fyv;
      const bool cj = (S && jj == jsc) || (N && jj == jec + 1);
      if (W && ii == isc && cj) vc_ = fx1v - fxv + fyv;
      if (E && ii == iec + 1 && cj) vc_ = fx1v - fxv - fy1v;
      return m.fC[oo2] + m.rarea_c[oo2] * vc_;
    };
    if (i <= iec + 1) {
      double vcv = vc0[o];
      if (i >= isc && i <= iec && j >= jsc && j <= jec + 1) {
        double tmp = dt2 * (u[o] - vcv * m.cosa_v[o2]) / m.sina_v[o2];
        if ((S && j == jsc) || (N && j == jec + 1)) tmp = dt2 * u[o];
        const double flux = vort(tmp > 0.0 ? i : i + 1, j);
        vcv = vcv - tmp * flux + m.rdyc[o2] * (ke[o - sj] - ke[o]);
      }
      vc[o] = vcv;
    }
    if (j <= jec + 1) {
      double ucv = uc0[o];
      if (i >= isc && i <= iec + 1 && j >= jsc && j <= jec) {
        double tmp = dt2 * (v[o] - ucv * m.cosa_u[o2]) / m.sina_u[o2];
        if ((W && i == isc) || (E && i == iec + 1)) tmp = dt2 * v[o];
        const double flux = vort(i, tmp > 0.0 ? j : j + 1);
        ucv = ucv + tmp * flux + m.rdxc[o2] * (ke[o - 1] - ke[o]);
      }
      uc[o] = ucv;
    }
  });
  return fv3::check_launch("fv3_c_sw");
}

}  // extern "C"
