// 1-D piecewise-parabolic "value advected through an interface" and cube-corner index remapping.
//   ppm_al_lt8 / ppm_blbr8 / ppm_flux_staged
//             <- compute_x_flux / compute_y_flux (fv3core/pace/fv3core/stencils/xppm.py:249-266, yppm.py mirror):
//                ord < 8: compute_al (:148-181) + get_flux (:64-71) with the monotonicity mask (:47-61);
//                ord == 8: compute_blbr_ord8plus (:249-262) = dm/al/blbr (:82-102) + bl_br_edges (:185-246)
//                          + pert_ppm_standard_constraint_fcn (ppm.py:22-36) + get_flux_ord8plus (:74-79)
//   corner_x / corner_y <- copy_corners_x/y_stencil_defn (stencils/pace/stencils/corners.py:307-425): instead of
//                copying the 3x3 cube-corner halo blocks in place before every sweep, reads are redirected.
#pragma once
#include "common.h"

namespace fv3 {

constexpr double PPM_C1 = -2.0 / 14.0, PPM_C2 = 11.0 / 14.0, PPM_C3 = 5.0 / 14.0;
constexpr double PPM_P1 = 7.0 / 12.0, PPM_P2 = -1.0 / 12.0;
constexpr double PPM_S11 = 11.0 / 14.0, PPM_S14 = 4.0 / 7.0, PPM_S15 = 3.0 / 14.0;

// basic_operations.sign (basic_operations.py:32-39): +|a| only for b > 0
// On the device the sign bit is set directly (one compare, one select, one logic op instead of two negations, a
// compare and two selects); the value is the same bit for bit, signed zeros and NaN operands included.
FV_HD double rsign(double a, double b) {
#if defined(__CUDA_ARCH__)
  const int hi = (__double2hiint(a) & 0x7fffffff) | (b > 0 ? 0 : (int)0x80000000);
  return __hiloint2double(hi, __double2loint(a));
#else
  return b > 0 ? fabs(a) : -fabs(a);
#endif
}

struct Edge1D {
  bool lo, hi;     // subdomain touches the low / high tile edge in this direction
  int start, end;  // first / last compute index in this direction (i_start, i_end of the reference)
};

// q(i): cell value at absolute index i along the sweep direction; dx(i): cell width (dxa or dya)
template <class Q, class DX>
FV_HD double ppm_al_lt8(Q q, DX dx, int i, const Edge1D &e) {
  if ((e.lo && i == e.start - 1) || (e.hi && i == e.end)) return PPM_C1 * q(i - 2) + PPM_C2 * q(i - 1) + PPM_C3 * q(i);
  if ((e.lo && i == e.start) || (e.hi && i == e.end + 1))
    return 0.5 * (((2.0 * dx(i - 1) + dx(i - 2)) * q(i - 1) - dx(i - 1) * q(i - 2)) / (dx(i - 2) + dx(i - 1)) +
                  ((2.0 * dx(i) + dx(i + 1)) * q(i) - dx(i) * q(i + 1)) / (dx(i) + dx(i + 1)));
  if ((e.lo && i == e.start + 1) || (e.hi && i == e.end + 2)) return PPM_C3 * q(i - 1) + PPM_C2 * q(i) + PPM_C1 * q(i + 1);
  return PPM_P1 * (q(i - 1) + q(i)) + PPM_P2 * (q(i - 2) + q(i + 1));
}

FV_HD double ppm_fx1(double c, double br_l, double b0_l, double bl_r, double b0_r) {
  return c > 0.0 ? (1.0 - c) * (br_l - c * b0_l) : (1.0 + c) * (bl_r + c * b0_r);
}

template <class Q>
FV_HD double ppm_dm8(Q q, int i) {
  const double xt = 0.25 * (q(i + 1) - q(i - 1));
  const double dqr = dmax(dmax(q(i), q(i - 1)), q(i + 1)) - q(i);
  const double dql = q(i) - dmin(dmin(q(i), q(i - 1)), q(i + 1));
  return rsign(dmin(dmin(fabs(xt), dqr), dql), xt);
}
template <class Q>
FV_HD double ppm_al8(Q q, int i) {
  return 0.5 * (q(i - 1) + q(i)) + 1.0 / 3.0 * (ppm_dm8(q, i - 1) - ppm_dm8(q, i));
}
template <class Q, class DX>
FV_HD double ppm_edge0(Q q, DX dx, int i, bool minmax) {  // xt_dxa_edge_0
  double xt = 0.5 * (((2.0 * dx(i) + dx(i - 1)) * q(i) - dx(i) * q(i - 1)) / (dx(i - 1) + dx(i)) +
                     ((2.0 * dx(i + 1) + dx(i + 2)) * q(i + 1) - dx(i + 1) * q(i + 2)) / (dx(i + 1) + dx(i + 2)));
  if (minmax) {
    const double mn = dmin(dmin(dmin(q(i - 1), q(i)), q(i + 1)), q(i + 2));
    const double mx = dmax(dmax(dmax(q(i - 1), q(i)), q(i + 1)), q(i + 2));
    xt = dmin(dmax(xt, mn), mx);
  }
  return xt;
}
template <class Q, class DX>
FV_HD double ppm_edge1(Q q, DX dx, int i, bool minmax) {  // xt_dxa_edge_1
  double xt = 0.5 * (((2.0 * dx(i - 1) + dx(i - 2)) * q(i - 1) - dx(i - 1) * q(i - 2)) / (dx(i - 2) + dx(i - 1)) +
                     ((2.0 * dx(i) + dx(i + 1)) * q(i) - dx(i) * q(i + 1)) / (dx(i) + dx(i + 1)));
  if (minmax) {
    const double mn = dmin(dmin(dmin(q(i - 2), q(i - 1)), q(i)), q(i + 1));
    const double mx = dmax(dmax(dmax(q(i - 2), q(i - 1)), q(i)), q(i + 1));
    xt = dmin(dmax(xt, mn), mx);
  }
  return xt;
}
// dm of cell c computed the way bl_br_edges does it for its neighbours (same formula as ppm_dm8)
template <class Q, class DX>
FV_HD void ppm_blbr8(Q q, DX dx, int i, const Edge1D &e, bool minmax, double &bl, double &br) {
  int kind = 0;
  if (e.lo && i >= e.start - 1 && i <= e.start + 1) kind = i - (e.start - 1) + 1;  // 1, 2, 3
  if (e.hi && i >= e.end - 1 && i <= e.end + 1) kind = i - (e.end - 1) + 4;        // 4, 5, 6
  if (kind == 0) {
    const double dm = ppm_dm8(q, i), xt = 2.0 * dm;
    bl = -1.0 * rsign(dmin(fabs(xt), fabs(ppm_al8(q, i) - q(i))), xt);
    br = rsign(dmin(fabs(xt), fabs(ppm_al8(q, i + 1) - q(i))), xt);
    return;
  }
  double xt_bl, xt_br;
  switch (kind) {
    case 1:  // i_start - 1
      xt_bl = PPM_S14 * ppm_dm8(q, i - 1) + PPM_S11 * (q(i - 1) - q(i)) + q(i);
      xt_br = ppm_edge0(q, dx, i, minmax);
      break;
    case 2:  // i_start
      xt_bl = ppm_edge1(q, dx, i, minmax);
      xt_br = PPM_S15 * q(i) + PPM_S11 * q(i + 1) - PPM_S14 * ppm_dm8(q, i + 1);
      break;
    case 3:  // i_start + 1
      xt_bl = PPM_S15 * q(i - 1) + PPM_S11 * q(i) - PPM_S14 * ppm_dm8(q, i);
      xt_br = ppm_al8(q, i + 1);
      break;
    case 4:  // i_end - 1
      xt_bl = ppm_al8(q, i);
      xt_br = PPM_S15 * q(i + 1) + PPM_S11 * q(i) + PPM_S14 * ppm_dm8(q, i);
      break;
    case 5:  // i_end
      xt_bl = PPM_S15 * q(i) + PPM_S11 * q(i - 1) + PPM_S14 * ppm_dm8(q, i - 1);
      xt_br = ppm_edge0(q, dx, i, minmax);
      break;
    default:  // i_end + 1
      xt_bl = ppm_edge1(q, dx, i, minmax);
      xt_br = PPM_S11 * (q(i + 1) - q(i)) - PPM_S14 * ppm_dm8(q, i + 1) + q(i);
      break;
  }
  double al = xt_bl - q(i), ar = xt_br - q(i);
  // pert_ppm_standard_constraint_fcn (ppm.py:22-36)
  if (al * ar < 0.0) {
    const double da1 = al - ar, da2 = da1 * da1, a6da = 3.0 * (al + ar) * da1;
    if (a6da < -da2)
      ar = -2.0 * al;
    else if (a6da > da2)
      al = -2.0 * ar;
  } else {
    al = 0.0;
    ar = 0.0;
  }
  bl = al;
  br = ar;
}

// ---- two-pass form used by the plane-resident kernels (sweep.h): pass 1 stores, per line, the edge values al
// (MORD < 8) or the limited slopes dm (MORD == 8) ONCE per face / cell; pass 2 builds the interface value from them.
// T: accessor of the staged values (al at face i / dm of cell i)
template <int MORD, class Q, class T, class DX>
FV_HD double ppm_flux_staged(Q q, T t, DX dx, double c, int i, const Edge1D &e) {
  if (MORD < 8) {
    const double al0 = t(i - 1), al1 = t(i), al2 = t(i + 1);
    const double ql = q(i - 1), qr = q(i);
    const double bl_l = al0 - ql, br_l = al1 - ql, b0_l = bl_l + br_l;
    const double bl_r = al1 - qr, br_r = al2 - qr, b0_r = bl_r + br_r;
    bool s_l, s_r;
    if (MORD == 5) {
      s_l = bl_l * br_l < 0;
      s_r = bl_r * br_r < 0;
    } else {
      s_l = (3.0 * fabs(b0_l)) < fabs(bl_l - br_l);
      s_r = (3.0 * fabs(b0_r)) < fabs(bl_r - br_r);
    }
    const double mask = (s_l || s_r) ? 1.0 : 0.0;
    const double fx1 = ppm_fx1(c, br_l, b0_l, bl_r, b0_r);
    return c > 0.0 ? ql + fx1 * mask : qr + fx1 * mask;
  }
  const bool pos = c > 0.0;
  const int cc = pos ? i - 1 : i;
  const double qc = q(cc);
  double bl, br;
  if ((e.lo && cc >= e.start - 1 && cc <= e.start + 1) || (e.hi && cc >= e.end - 1 && cc <= e.end + 1)) {
    ppm_blbr8(q, dx, cc, e, true, bl, br);
  } else {
    const double dm0 = t(cc), xt = 2.0 * dm0;
    const double alc = 0.5 * (q(cc - 1) + qc) + 1.0 / 3.0 * (t(cc - 1) - dm0);
    const double alr = 0.5 * (qc + q(cc + 1)) + 1.0 / 3.0 * (dm0 - t(cc + 1));
    bl = -1.0 * rsign(dmin(fabs(xt), fabs(alc - qc)), xt);
    br = rsign(dmin(fabs(xt), fabs(alr - qc)), xt);
  }
  const double b0 = bl + br;
  return pos ? qc + (1.0 - c) * (br - c * b0) : qc + (1.0 + c) * (bl + c * b0);
}

// copy_corners_x: a cell in a cube-corner halo block at outward distances (a, b) reads the cell at
// x outward distance b, y inward depth a.
FV_HD void corner_x(const fv3_geom &g, int s, int &i, int &j) {
  const int isc = g.halo, iec = g.halo + g.nx - 1, jsc = g.halo, jec = g.halo + g.ny - 1;
  const bool xo_w = i < isc, xo_e = i > iec, yo_s = j < jsc, yo_n = j > jec;
  if (!((xo_w || xo_e) && (yo_s || yo_n))) return;
  if (!((xo_w ? on_west(g, s) : on_east(g, s)) && (yo_s ? on_south(g, s) : on_north(g, s)))) return;
  const int a = xo_w ? isc - i : i - iec, b = yo_s ? jsc - j : j - jec;
  i = xo_w ? isc - b : iec + b;
  j = yo_s ? jsc + a - 1 : jec - a + 1;
}
// copy_corners_y: reads the cell at x inward depth b, y outward distance a.
FV_HD void corner_y(const fv3_geom &g, int s, int &i, int &j) {
  const int isc = g.halo, iec = g.halo + g.nx - 1, jsc = g.halo, jec = g.halo + g.ny - 1;
  const bool xo_w = i < isc, xo_e = i > iec, yo_s = j < jsc, yo_n = j > jec;
  if (!((xo_w || xo_e) && (yo_s || yo_n))) return;
  if (!((xo_w ? on_west(g, s) : on_east(g, s)) && (yo_s ? on_south(g, s) : on_north(g, s)))) return;
  const int a = xo_w ? isc - i : i - iec, b = yo_s ? jsc - j : j - jec;
  i = xo_w ? isc + b - 1 : iec - b + 1;
  j = yo_s ? jsc - a : jec + a;
}

}  // namespace fv3
