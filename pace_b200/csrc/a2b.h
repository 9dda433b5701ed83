// A-grid -> B-grid (cell corner) 4th-order interpolation.
//   a2b_plane <- AGrid2BGridFourthOrder.__call__ (fv3core/pace/fv3core/stencils/a2b_ord4.py:673-761):
//                corner extrapolation (:37-273; great-circle weights precomputed on the host, fv3_grid.a2b_w),
//                tile-edge formulas qout_x_edge / qout_y_edge (:286-311), ppm_volume_mean_x/y (:416-450) and
//                a2b_interpolation (:453-481).
#pragma once
#include "common.h"
#include "plane.h"

namespace fv3 {

struct A2B {
  static constexpr double c1 = 2.0 / 3.0, c2 = -1.0 / 6.0, b1 = 7.0 / 12.0, b2 = -1.0 / 12.0, a1 = 9.0 / 16.0,
                          a2 = -1.0 / 16.0;
};

// Q: qin(i, j) accessor
template <class Q>
FV_HD double a2b_qx(const fv3_geom &g, const fv3_grid &m, int s, Q q, int i, int j) {
  const int isc = g.halo, iec = g.halo + g.nx - 1;
  auto dxa = [&](int ii) { return m.dxa[O2(s, ii, j)]; };
  if (on_west(g, s)) {
    if (i == isc) {
      const double g_in = dxa(i + 1) / dxa(i), g_ou = dxa(i - 2) / dxa(i - 1);
      return 0.5 * (((2.0 + g_in) * q(i, j) - q(i + 1, j)) / (1.0 + g_in) + ((2.0 + g_ou) * q(i - 1, j) - q(i - 2, j)) / (1.0 + g_ou));
    }
    if (i == isc + 1) {
      const double g_in = dxa(i) / dxa(i - 1), g_ou = dxa(i - 3) / dxa(i - 2);
      const double qxleft = 0.5 * (((2.0 + g_in) * q(i - 1, j) - q(i, j)) / (1.0 + g_in) + ((2.0 + g_ou) * q(i - 2, j) - q(i - 3, j)) / (1.0 + g_ou));
      const double qxright = A2B::b2 * (q(i - 1, j) + q(i + 2, j)) + A2B::b1 * (q(i, j) + q(i + 1, j));
      return (3.0 * (g_in * q(i - 1, j) + q(i, j)) - (g_in * qxleft + qxright)) / (2.0 + 2.0 * g_in);
    }
  }
  if (on_east(g, s)) {
    if (i == iec + 1) {
      const double g_in = dxa(i - 2) / dxa(i - 1), g_ou = dxa(i + 1) / dxa(i);
      return 0.5 * (((2.0 + g_in) * q(i - 1, j) - q(i - 2, j)) / (1.0 + g_in) + ((2.0 + g_ou) * q(i, j) - q(i + 1, j)) / (1.0 + g_ou));
    }
    if (i == iec) {
      const double g_in = dxa(i - 1) / dxa(i), g_ou = dxa(i + 2) / dxa(i + 1);
      const double qxright = 0.5 * (((2.0 + g_in) * q(i, j) - q(i - 1, j)) / (1.0 + g_in) + ((2.0 + g_ou) * q(i + 1, j) - q(i + 2, j)) / (1.0 + g_ou));
      const double qxleft = A2B::b2 * (q(i - 3, j) + q(i, j)) + A2B::b1 * (q(i - 2, j) + q(i - 1, j));
      return (3.0 * (q(i - 1, j) + g_in * q(i, j)) - (g_in * qxright + qxleft)) / (2.0 + 2.0 * g_in);
    }
  }
  return A2B::b2 * (q(i - 2, j) + q(i + 1, j)) + A2B::b1 * (q(i - 1, j) + q(i, j));
}

template <class Q>
FV_HD double a2b_qy(const fv3_geom &g, const fv3_grid &m, int s, Q q, int i, int j) {
  const int jsc = g.halo, jec = g.halo + g.ny - 1;
  auto dya = [&](int jj) { return m.dya[O2(s, i, jj)]; };
  if (on_south(g, s)) {
    if (j == jsc) {
      const double g_in = dya(j + 1) / dya(j), g_ou = dya(j - 2) / dya(j - 1);
      return 0.5 * (((2.0 + g_in) * q(i, j) - q(i, j + 1)) / (1.0 + g_in) + ((2.0 + g_ou) * q(i, j - 1) - q(i, j - 2)) / (1.0 + g_ou));
    }
    if (j == jsc + 1) {
      const double g_in = dya(j) / dya(j - 1), g_ou = dya(j - 3) / dya(j - 2);
      const double lower = 0.5 * (((2.0 + g_in) * q(i, j - 1) - q(i, j)) / (1.0 + g_in) + ((2.0 + g_ou) * q(i, j - 2) - q(i, j - 3)) / (1.0 + g_ou));
      const double upper = A2B::b2 * (q(i, j - 1) + q(i, j + 2)) + A2B::b1 * (q(i, j) + q(i, j + 1));
      return (3.0 * (g_in * q(i, j - 1) + q(i, j)) - (g_in * lower + upper)) / (2.0 + 2.0 * g_in);
    }
  }
  if (on_north(g, s)) {
    if (j == jec + 1) {
      const double g_in = dya(j - 2) / dya(j - 1), g_ou = dya(j + 1) / dya(j);
      return 0.5 * (((2.0 + g_in) * q(i, j - 1) - q(i, j - 2)) / (1.0 + g_in) + ((2.0 + g_ou) * q(i, j) - q(i, j + 1)) / (1.0 + g_ou));
    }
    if (j == jec) {
      const double g_in = dya(j - 1) / dya(j), g_ou = dya(j + 2) / dya(j + 1);
      const double lower = A2B::b2 * (q(i, j - 3) + q(i, j)) + A2B::b1 * (q(i, j - 2) + q(i, j - 1));
      const double upper = 0.5 * (((2.0 + g_in) * q(i, j) - q(i, j - 1)) / (1.0 + g_in) + ((2.0 + g_ou) * q(i, j + 1) - q(i, j + 2)) / (1.0 + g_ou));
      return (3.0 * (q(i, j - 1) + g_in * q(i, j)) - (g_in * upper + lower)) / (2.0 + 2.0 * g_in);
    }
  }
  return A2B::b2 * (q(i, j - 2) + q(i, j + 1)) + A2B::b1 * (q(i, j - 1) + q(i, j));
}

// value on a tile edge / tile corner point (the reference's tmp_qout_edges)
template <class Q>
FV_HD double a2b_edge_value(const fv3_geom &g, const fv3_grid &m, int s, Q q, int i, int j) {
  const int isc = g.halo, iec = g.halo + g.nx - 1, jsc = g.halo, jec = g.halo + g.ny - 1;
  const bool iw = on_west(g, s) && i == isc, ie = on_east(g, s) && i == iec + 1;
  const bool js = on_south(g, s) && j == jsc, jn = on_north(g, s) && j == jec + 1;
  if ((iw || ie) && (js || jn)) {
    // extrap_corner (a2b_ord4.py:43-56) with host-computed weights w = x1 / (x2 - x1)
    const int c = (js ? 0 : 2) + (iw ? 0 : 1);  // 0 sw, 1 se, 2 nw, 3 ne
    const double *w = m.a2b_w + ((int64_t)s * 4 + c) * 3;
    const int di = iw ? 1 : -1, dj = js ? 1 : -1;  // direction pointing into the tile
    const int i0 = iw ? i : i - 1, j0 = js ? j : j - 1;  // first interior cell next to the corner point
    // the three "arms": into this tile, across the x edge, across the y edge
    const double qa1 = q(i0, j0), qb1 = q(i0 + di, j0 + dj);
    const double qa2 = q(i0 - di, j0), qb2 = q(i0 - 2 * di, j0 + dj);
    const double qa3 = q(i0, j0 - dj), qb3 = q(i0 + di, j0 - 2 * dj);
    const double e1 = qa1 + w[0] * (qa1 - qb1), e2 = qa2 + w[1] * (qa2 - qb2), e3 = qa3 + w[2] * (qa3 - qb3);
    // arms: e1 into this tile, e2 across the x edge, e3 across the y edge; the reference sums them in this
    // order at every corner except its "_nw_corner" stencil at (iec+1, jsc) (a2b_ord4.py:59-273)
    const double sum = (c == 1) ? (e1 + e3 + e2) : (e1 + e2 + e3);
    return sum * (1.0 / 3.0);
  }
  if (iw || ie) {
    auto q2 = [&](int jj) {
      return (q(i - 1, jj) * m.dxa[O2(s, i, jj)] + q(i, jj) * m.dxa[O2(s, i - 1, jj)]) / (m.dxa[O2(s, i - 1, jj)] + m.dxa[O2(s, i, jj)]);
    };
    const double ew = (iw ? m.edge_w : m.edge_e)[O2(s, 0, j)];
    return ew * q2(j - 1) + (1.0 - ew) * q2(j);
  }
  auto q1 = [&](int ii) {
    return (q(ii, j - 1) * m.dya[O2(s, ii, j)] + q(ii, j) * m.dya[O2(s, ii, j - 1)]) / (m.dya[O2(s, ii, j - 1)] + m.dya[O2(s, ii, j)]);
  };
  const double es = (js ? m.edge_s : m.edge_n)[O2(s, i, 0)];
  return es * q1(i - 1) + (1.0 - es) * q1(i);
}

// ---- plane-resident form (plane.h): the reference's three temporaries qx, qy, qout_edges are built ONCE per strip
// in shared memory instead of being re-derived inside every thread.
//   SQ: qin plane   QX / QY: ppm_volume_mean_x / _y   OUT: B-grid result (tile-edge values first, then the rest)
// qin is a global pointer to the start of the (s, k) plane; the arrays are shared planes of the block (Block::plane).
// On return OUT holds the corner rows [ja, jb] of the strip — or, with `gout` (start of a global (s, k) plane), the
// rows the strip owns ([ja, jtop]) are written there directly and OUT keeps only the tile-edge values.
// Two block-wide passes: (1) qx, qy and the tile-edge values, all functions of qin only; (2) the interpolation.
template <class B>
FV_DEV void a2b_plane(const fv3_geom &g, const fv3_grid &m, int s, const B &b, const double *qin, double *SQ, double *QX,
                     double *QY, double *OUT, bool loaded = false, double *gout = nullptr) {
  const int sj = g.sj, h = g.halo;
  const int isc = h, iec = h + g.nx - 1, jsc = h, jec = h + g.ny - 1;
  const bool W = on_west(g, s), E = on_east(g, s), S = on_south(g, s), N = on_north(g, s);
  const int nwi = g.ni - 1, nwj = g.nj - 1;
  const int ja = b.ja, jb = b.jb;
  // loaded: SQ already holds qin on every resident row (the caller produced it in shared memory)
  if (!loaded) b.rect(0, nwi, b.lo(0, h), b.hi(nwj, h), [&](int i, int j) { SQ[j * sj + i] = qin[j * sj + i]; });
  auto q = [&](int ii, int jj) { return SQ[jj * sj + ii]; };
  // qx on corner columns isc..iec+1, rows ja-2..jb+1; qy on rows ja..jb, columns isc-2..iec+2
  // (a last strip of ONE row on a north tile edge: its row jec reads qx three rows down, a2b_ord4.py:286-311)
  const int xj0 = b.lo(jsc - 2, (N && jb - ja < 2) ? 3 : 2), xj1 = b.hi(jec + 3, 2), yj0 = ja, yj1 = jb + 1;
  const int nxc = g.nx + 1, nxw = g.nx + 4, nxr = xj1 - xj0, nyr = yj1 - yj0;
  // The two columns (rows) either side of a tile edge use the one-sided formulas (divides, metric loads): they are
  // separate, densely packed tasks of the same phase instead of a few slow lanes in every warp of the bulk pass.
  const int nbx = nxc * nxr, nby = nxw * nyr, nex = 4 * nxr, ney = 4 * nxw;
  // tile-edge values, one corner row beyond the strip (the rows next to a tile edge read them): the points of the west /
  // east corner columns and of the south / north corner rows are tasks of the same pass
  const int ej0 = b.lo(jsc, 1), ej1 = b.hi(jec + 2, 2), ner = ej1 > ej0 ? ej1 - ej0 : 0;
  const int nev = (W || E || S || N) ? 2 * ner + 2 * nxc : 0;
  const float inv_c = 1.0f / (float)nxc, inv_w = 1.0f / (float)nxw;
  b.par(nbx + nby + nex + ney + nev, [&](int t) {
    if (t < nbx) {
      const int jr = row_of(t, nxc, inv_c), i = isc + (t - jr * nxc), j = xj0 + jr;
      if (i <= isc + 1 || i >= iec) return;
      QX[j * sj + i] = A2B::b2 * (q(i - 2, j) + q(i + 1, j)) + A2B::b1 * (q(i - 1, j) + q(i, j));
    } else if (t < nbx + nby) {
      const int t2 = t - nbx;
      const int jr = row_of(t2, nxw, inv_w), i = isc - 2 + (t2 - jr * nxw), j = yj0 + jr;
      if (j <= jsc + 1 || j >= jec) return;
      QY[j * sj + i] = A2B::b2 * (q(i, j - 2) + q(i, j + 1)) + A2B::b1 * (q(i, j - 1) + q(i, j));
    } else if (t < nbx + nby + nex) {
      const int t2 = t - nbx - nby;
      const int c = t2 & 3, j = xj0 + (t2 >> 2), i = c < 2 ? isc + c : iec + (c - 2);
      if (c >= 2 && i <= isc + 1) return;  // tiny domains: column already done as a west column
      QX[j * sj + i] = a2b_qx(g, m, s, q, i, j);
    } else if (t < nbx + nby + nex + ney) {
      const int t2 = t - nbx - nby - nex;
      const int c = row_of(t2, nxw, inv_w), i = isc - 2 + (t2 - c * nxw), j = c < 2 ? jsc + c : jec + (c - 2);
      if (c >= 2 && j <= jsc + 1) return;
      if (j < yj0 || j >= yj1) return;
      QY[j * sj + i] = a2b_qy(g, m, s, q, i, j);
    } else {
      int t2 = t - nbx - nby - nex - ney, i, j;
      if (t2 < 2 * ner) {  // west, then east corner column
        const bool east = t2 >= ner;
        if (east ? !E : !W) return;
        i = east ? iec + 1 : isc;
        j = ej0 + (east ? t2 - ner : t2);
      } else {  // south, then north corner row
        t2 -= 2 * ner;
        const bool north = t2 >= nxc;
        if (north ? !N : !S) return;
        j = north ? jec + 1 : jsc;
        if (j < ej0 || j >= ej1) return;
        i = isc + (north ? t2 - nxc : t2);
      }
      const double v = a2b_edge_value(g, m, s, q, i, j);
      OUT[j * sj + i] = v;  // a tile-corner point is visited by a column and a row task: same value
      if (gout && j >= ja && j <= b.jtop()) gout[j * sj + i] = v;
    }
  });
  b.rect(isc, iec + 2, ja, (gout ? b.jtop() : jb) + 1, [&](int i, int j) {
    if ((W && i == isc) || (E && i == iec + 1) || (S && j == jsc) || (N && j == jec + 1)) return;
    auto qx = [&](int jj) { return QX[jj * sj + i]; };
    auto qy = [&](int ii) { return QY[j * sj + ii]; };
    double qxx, qyy;
    if (S && j == jsc + 1) {
      const double upper = A2B::a2 * (qx(j - 1) + qx(j + 2)) + A2B::a1 * (qx(j) + qx(j + 1));
      qxx = A2B::c1 * (qx(j - 1) + qx(j)) + A2B::c2 * (OUT[(j - 1) * sj + i] + upper);
    } else if (N && j == jec) {
      const double lower = A2B::a2 * (qx(j - 3) + qx(j)) + A2B::a1 * (qx(j - 2) + qx(j - 1));
      qxx = A2B::c1 * (qx(j - 1) + qx(j)) + A2B::c2 * (OUT[(j + 1) * sj + i] + lower);
    } else {
      qxx = A2B::a2 * (qx(j - 2) + qx(j + 1)) + A2B::a1 * (qx(j - 1) + qx(j));
    }
    if (W && i == isc + 1) {
      const double right = A2B::a2 * (qy(i - 1) + qy(i + 2)) + A2B::a1 * (qy(i) + qy(i + 1));
      qyy = A2B::c1 * (qy(i - 1) + qy(i)) + A2B::c2 * (OUT[j * sj + i - 1] + right);
    } else if (E && i == iec) {
      const double left = A2B::a2 * (qy(i - 3) + qy(i)) + A2B::a1 * (qy(i - 2) + qy(i - 1));
      qyy = A2B::c1 * (qy(i - 1) + qy(i)) + A2B::c2 * (OUT[j * sj + i + 1] + left);
    } else {
      qyy = A2B::a2 * (qy(i - 2) + qy(i + 1)) + A2B::a1 * (qy(i - 1) + qy(i));
    }
    if (gout)
      gout[j * sj + i] = 0.5 * (qxx + qyy);
    else
      OUT[j * sj + i] = 0.5 * (qxx + qyy);
  });
}

}  // namespace fv3
