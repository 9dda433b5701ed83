// One PPM sweep of a shared-memory plane (xppm.py / yppm.py via ppm.h) as line marches in registers.
//
// A task is a run of NF consecutive faces of one line.  Its thread walks the run with a sliding register window:
// per face ONE shared-memory load (the next cell value), one new edge value al (hord 5/6) or limited slope dm and
// cell parabola bl/br (hord 8), one interface value — no staging plane, no intermediate barrier, index arithmetic
// once per task.  Runs along i start NF doubles apart with NF odd, so the 32 lanes of a warp hit distinct banks;
// runs along j put consecutive columns on consecutive lanes.
// On cube-tile edges the 3 faces either side whose parabola uses the one-sided edge formulas are left out of the
// march and evaluated by extra tasks of the same phase with the edge-aware one-pass forms (ppm_flux_lt8 /
// ppm_flux8_upwind): same expressions, so every face gets the value the reference computes, exactly once.
#pragma once
#include "plane.h"
#include "ppm.h"

namespace fv3 {

// faces per task: the odd run length that minimises (trips over the block) x (steps per run)
FV_HD int sweep_run_length(int nlines, int nfaces, int nthreads) {
  int best = 3, best_cost = 1 << 30;
  for (int nf = 3; nf <= 15; nf += 2) {
    const int tasks = nlines * ((nfaces + nf - 1) / nf);
    const int cost = ((tasks + nthreads - 1) / nthreads) * (nf + 8);
    if (cost < best_cost) {
      best_cost = cost;
      best = nf;
    }
  }
  return best;
}

// Qs: values.  XDIR: sweep along i (stride 1) on lines j in [l0, l0+nl), else along j (stride sj) on lines i in
// [l0, l0+nl).  cg: Courant numbers, dxg: cell widths (global planes, same offsets).  Interface values are produced
// for the faces [f0, f1] of every line (the whole line: e.start .. e.end + 1; a strip sweeping along j passes its own
// face rows).  fin(p, value): what to do with the value at plane offset p (called exactly once per face).
template <int MORD, bool XDIR, class Fin>
FV_DEV void ppm_sweep(const Block &b, const double *Qs, int sj, const double *cg, const double *dxg, const Edge1D &e,
                      int l0, int nl, int f0, int f1, Fin fin) {
  const int st = XDIR ? 1 : sj, ls = XDIR ? sj : 1;
  if (nl <= 0 || f1 < f0) return;  // uniform over the block
  const int n = f1 - f0 + 1;
#ifdef FV3_HOSTSIM
  const int NF = 9;
#else
  const int NF = sweep_run_length(nl, n, (int)blockDim.x);
#endif
  const int nch = (n + NF - 1) / NF, ntask = nl * nch;
  // faces evaluated by the edge tasks
  const int lo0 = e.lo ? e.start : 1 << 30, lo1 = e.lo ? e.start + 2 : -(1 << 30);
  const int hi0 = e.hi ? e.end - 1 : 1 << 30, hi1 = e.hi ? e.end + 1 : -(1 << 30);
  const int nfix = (e.lo || e.hi) ? 6 * nl : 0;
  b.par(ntask + nfix, [&](int t) {
    if (t >= ntask) {
      const int t2 = t - ntask, l = l0 + t2 / 6, r = t2 % 6;
      if (r < 3 ? !e.lo : !e.hi) return;
      const int f = r < 3 ? e.start + r : e.end - 1 + (r - 3);
      if (r >= 3 && e.lo && f <= e.start + 2) return;  // tiny domains: already done by the low-edge tasks
      if (f < f0 || f > f1) return;
      const int p = f * st + l * ls;
      auto q = [&](int ii) { return Qs[ii * st + l * ls]; };
      auto dx = [&](int ii) { return dxg[ii * st + l * ls]; };
      if (MORD < 8)
        fin(p, ppm_flux_lt8(MORD, q, dx, cg[p], f, e));
      else
        fin(p, ppm_flux8_upwind(q, dx, cg[p], f, e, true));
      return;
    }
    int l, ch;
    if (XDIR) {
      l = t / nch;
      ch = t - l * nch;
    } else {
      ch = t / nl;
      l = t - ch * nl;
    }
    l += l0;
    const int fa = f0 + ch * NF, fb = fa + NF - 1 < f1 ? fa + NF - 1 : f1;
    int p = fa * st + l * ls;
    const double *qp = Qs + p;
    double cn = FV_LDG(cg + p);
    if (MORD < 8) {
      // window: q0..q3 = q[f-2..f+1], alm = al[f-1], al0 = al[f]
      double q0 = qp[-2 * st], q1 = qp[-st], q2 = qp[0], q3 = qp[st];
      double alm = PPM_P1 * (q0 + q1) + PPM_P2 * (qp[-3 * st] + q2);
      double al0 = PPM_P1 * (q1 + q2) + PPM_P2 * (q0 + q3);
      for (int f = fa; f <= fb; ++f) {
        const double q4 = qp[2 * st];
        const double c = cn;
        if (f < fb) cn = FV_LDG(cg + p + st);
        const double al2 = PPM_P1 * (q2 + q3) + PPM_P2 * (q1 + q4);
        if (!((f >= lo0 && f <= lo1) || (f >= hi0 && f <= hi1))) {
          const double ql = q1, qr = q2;
          const double bl_l = alm - ql, br_l = al0 - ql, b0_l = bl_l + br_l;
          const double bl_r = al0 - qr, br_r = al2 - qr, b0_r = bl_r + br_r;
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
        }
        q0 = q1;
        q1 = q2;
        q2 = q3;
        q3 = q4;
        alm = al0;
        al0 = al2;
        p += st;
        qp += st;
      }
    } else {
      // limited slope of a cell from its value and its two neighbours (xppm.py:85-89)
      auto dmf = [](double qm, double q0, double qq) {
        const double xt = 0.25 * (qq - qm);
        const double dqr = dmax(dmax(q0, qm), qq) - q0;
        const double dql = q0 - dmin(dmin(q0, qm), qq);
        return rsign(dmin(dmin(fabs(xt), dqr), dql), xt);
      };
      // window: qa, qb, qc = q[f-1], q[f], q[f+1]; dma, dmb = dm[f-1], dm[f]; (ql, bll, brl) = parabola of cell f-1
      double qa = qp[-st], qb = qp[0], qc = qp[st];
      const double qm2 = qp[-2 * st];
      const double dmm = dmf(qp[-3 * st], qm2, qa);  // dm[f-2]
      double dma = dmf(qm2, qa, qb), dmb = dmf(qa, qb, qc);
      double ql = qa, bll, brl;
      {
        const double xt = 2.0 * dma;
        const double alc = 0.5 * (qm2 + qa) + 1.0 / 3.0 * (dmm - dma);
        const double alr = 0.5 * (qa + qb) + 1.0 / 3.0 * (dma - dmb);
        bll = -1.0 * rsign(dmin(fabs(xt), fabs(alc - qa)), xt);
        brl = rsign(dmin(fabs(xt), fabs(alr - qa)), xt);
      }
      for (int f = fa; f <= fb; ++f) {
        const double qd = qp[2 * st];
        const double c = cn;
        if (f < fb) cn = FV_LDG(cg + p + st);
        const double dmc = dmf(qb, qc, qd);  // dm[f+1]
        // parabola of cell f
        const double xt = 2.0 * dmb;
        const double alc = 0.5 * (qa + qb) + 1.0 / 3.0 * (dma - dmb);
        const double alr = 0.5 * (qb + qc) + 1.0 / 3.0 * (dmb - dmc);
        const double blr = -1.0 * rsign(dmin(fabs(xt), fabs(alc - qb)), xt);
        const double brr = rsign(dmin(fabs(xt), fabs(alr - qb)), xt);
        if (!((f >= lo0 && f <= lo1) || (f >= hi0 && f <= hi1))) {
          const bool pos = c > 0.0;
          const double q0 = pos ? ql : qb, bl = pos ? bll : blr, br = pos ? brl : brr;
          const double b0 = bl + br;
          fin(p, pos ? q0 + (1.0 - c) * (br - c * b0) : q0 + (1.0 + c) * (bl + c * b0));
        }
        ql = qb;
        bll = blr;
        brl = brr;
        qa = qb;
        qb = qc;
        qc = qd;
        dma = dmb;
        dmb = dmc;
        p += st;
        qp += st;
      }
    }
  });
}

}  // namespace fv3
