// One PPM sweep of a shared-memory plane (xppm.py / yppm.py via ppm.h), organised for instruction count:
//   S1   stage the edge values al (hord 5/6) or limited slopes dm (hord 8) of every line with the branch-free
//        interior formula, addressing neighbours as constant offsets from one pointer per point;
//   S1b  re-evaluate the 3 faces next to a cube-tile edge with the edge formulas (hord 5/6 only, tile-edge CTAs only);
//   S2   interface value of every face from the staged line, interior formula, again branch-free;
//   S2b  hord 8 only: redo the <= 3 faces per tile edge whose upwind cell uses the one-sided bl/br formulas.
// Results are bit-identical to ppm_flux_staged / ppm_flux (same expressions, same order).
#pragma once
#include "plane.h"
#include "ppm.h"

namespace fv3 {

// Qs: values, Ts: staging plane (same layout).  XDIR: sweep along i (stride 1) on lines j in [l0, l0+nl), else along j
// (stride sj) on lines i in [l0, l0+nl).  cg: Courant numbers, dxg: cell widths (global planes, same offsets).
// Interface values are produced for the faces [f0, f1] of every line (the whole line: e.start .. e.end + 1; a strip
// sweeping along j passes its own face rows).  fin(p, value): what to do with the value at plane offset p.
template <int MORD, bool XDIR, class Fin>
FV_DEV void ppm_sweep(const Block &b, const double *Qs, double *Ts, int sj, const double *cg, const double *dxg,
                      const Edge1D &e, int l0, int nl, int f0, int f1, Fin fin) {
  const int st = XDIR ? 1 : sj, ls = XDIR ? sj : 1;
  const int st2 = 2 * st;
  const int start = e.start;
  if (nl <= 0 || f1 < f0) return;  // uniform over the block
  // staged values: al at faces f0-1 .. f1+1 (hord 5/6), dm of cells f0-2 .. f1+1 (hord 8)
  const int st0 = MORD < 8 ? -1 : -2, stn = MORD < 8 ? 3 : 4;
  const int n = f1 - f0 + 1;
  const int s1w = XDIR ? n + stn - 1 : nl, s1h = XDIR ? nl : n + stn - 1;
  b.par2(s1w, s1h, [&](int ir, int jr) {
    const int f = f0 + st0 + (XDIR ? ir : jr), l = l0 + (XDIR ? jr : ir);
    const double *qp = Qs + f * st + l * ls;
    if (MORD < 8) {
      Ts[f * st + l * ls] = PPM_P1 * (qp[-st] + qp[0]) + PPM_P2 * (qp[-st2] + qp[st]);
    } else {
      const double q0 = qp[0], qm = qp[-st], qq = qp[st];
      const double xt = 0.25 * (qq - qm);
      const double dqr = dmax(dmax(q0, qm), qq) - q0;
      const double dql = q0 - dmin(dmin(q0, qm), qq);
      Ts[f * st + l * ls] = rsign(dmin(dmin(fabs(xt), dqr), dql), xt);
    }
  });
  if (MORD < 8 && (e.lo || e.hi)) {
    b.par(6 * nl, [&](int t) {
      const int l = l0 + t / 6, r = t % 6;
      if (r < 3 ? !e.lo : !e.hi) return;
      const int f = r < 3 ? start - 1 + r : e.end + (r - 3);
      if (f < f0 - 1 || f > f1 + 1) return;
      auto q = [&](int ii) { return Qs[ii * st + l * ls]; };
      auto dx = [&](int ii) { return dxg[ii * st + l * ls]; };
      Ts[f * st + l * ls] = ppm_al_lt8(q, dx, f, e);
    });
  }
  const int s2w = XDIR ? n : nl, s2h = XDIR ? nl : n;
  // hord 8: faces redone by S2b are skipped here, so that fin() runs exactly once per face
  const int lo_lim = (MORD >= 8 && e.lo) ? start + 2 : start - 1, hi_lim = (MORD >= 8 && e.hi) ? e.end - 1 : e.end + 2;
  b.par2(s2w, s2h, [&](int ir, int jr) {
    const int f = f0 + (XDIR ? ir : jr), l = l0 + (XDIR ? jr : ir);
    if (f <= lo_lim || f >= hi_lim) return;
    const int p = f * st + l * ls;
    const double c = FV_LDG(cg + p);
    const double *qp = Qs + p, *tp = Ts + p;
    if (MORD < 8) {
      const double al0 = tp[-st], al1 = tp[0], al2 = tp[st];
      const double ql = qp[-st], qr = qp[0];
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
      fin(p, c > 0.0 ? ql + fx1 * mask : qr + fx1 * mask);
    } else {
      const bool pos = c > 0.0;
      const int oc = pos ? -st : 0;  // upwind cell relative to the face
      const double *qc = qp + oc, *tc = tp + oc;
      const double q0 = qc[0], dm0 = tc[0], xt = 2.0 * dm0;
      const double alc = 0.5 * (qc[-st] + q0) + 1.0 / 3.0 * (tc[-st] - dm0);
      const double alr = 0.5 * (q0 + qc[st]) + 1.0 / 3.0 * (dm0 - tc[st]);
      const double bl = -1.0 * rsign(dmin(fabs(xt), fabs(alc - q0)), xt);
      const double br = rsign(dmin(fabs(xt), fabs(alr - q0)), xt);
      const double b0 = bl + br;
      fin(p, pos ? q0 + (1.0 - c) * (br - c * b0) : q0 + (1.0 + c) * (bl + c * b0));
    }
  });
  if (MORD >= 8 && (e.lo || e.hi)) {
    // faces whose upwind cell can be one of the 3 cells either side of a tile edge: start..start+2, end-1..end+1
    b.par(6 * nl, [&](int t) {
      const int l = l0 + t / 6, r = t % 6;
      if (r < 3 ? !e.lo : !e.hi) return;
      const int f = r < 3 ? start + r : e.end - 1 + (r - 3);
      if (r >= 3 && e.lo && f <= start + 2) return;  // tiny domains: already done by the low-edge pass
      if (f < f0 || f > f1) return;
      const int p = f * st + l * ls;
      auto q = [&](int ii) { return Qs[ii * st + l * ls]; };
      auto tt = [&](int ii) { return Ts[ii * st + l * ls]; };
      auto dx = [&](int ii) { return dxg[ii * st + l * ls]; };
      fin(p, ppm_flux_staged<8>(q, tt, dx, cg[p], f, e));
    });
  }
}

}  // namespace fv3
