// One PPM sweep of a shared-memory plane (xppm.py / yppm.py via ppm.h), organised for instruction count and
// memory-level parallelism:
//   * a TASK is a run of 4 consecutive faces of one line.  The thread that owns it first issues every load of the
//     task — the 9-cell window of the line (x sweeps: five 128-bit shared loads; y sweeps: nine 64-bit loads,
//     conflict-free because consecutive lanes own consecutive columns) and the 4 Courant numbers — and then evaluates
//     the edge values al (hord 5/6) or limited slopes dm (hord 8) and the 4 interface values entirely in registers.
//     No staging plane, ONE barrier per sweep, one index decode per 4 faces.
//   * the <= 3 faces either side of a cube-tile edge whose stencil touches the one-sided edge formulas are skipped by
//     the bulk tasks and evaluated by separate, densely packed edge tasks of the same pass (tile-edge CTAs only).
// Results are bit-identical to ppm_flux_staged / ppm_flux (same expressions, same order).
#pragma once
#include "plane.h"
#include "ppm.h"

namespace fv3 {

constexpr int SWEEP_R = 4;  // faces per task

#ifdef FV3_HOSTSIM
FV_HD void ld_pair(const double *p, double &a, double &b) {
  a = p[0];
  b = p[1];
}
#else
__device__ __forceinline__ void ld_pair(const double *p, double &a, double &b) {
  const double2 v = *reinterpret_cast<const double2 *>(p);  // LDS.128 / LDG.128 (p is 16-byte aligned)
  a = v.x;
  b = v.y;
}
#endif

// interface value of one face from the register window: ql / qr = cells left / right of the face,
// al0..al2 = edge values at faces f-1, f, f+1 (hord 5/6)
template <int MORD>
FV_HD double ppm_face_lt8(double c, double ql, double qr, double al0, double al1, double al2) {
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
FV_HD double ppm_dm8v(double qm, double q0, double qq) {
  const double xt = 0.25 * (qq - qm);
  const double dqr = dmax(dmax(q0, qm), qq) - q0;
  const double dql = q0 - dmin(dmin(q0, qm), qq);
  return rsign(dmin(dmin(fabs(xt), dqr), dql), xt);
}
// hord 8: qm, q0, qq = upwind cell and its neighbours, dmm, dm0, dmp their limited slopes
FV_HD double ppm_face_8(double c, bool pos, double qm, double q0, double qq, double dmm, double dm0, double dmp) {
  const double xt = 2.0 * dm0;
  const double alc = 0.5 * (qm + q0) + 1.0 / 3.0 * (dmm - dm0);
  const double alr = 0.5 * (q0 + qq) + 1.0 / 3.0 * (dm0 - dmp);
  const double bl = -1.0 * rsign(dmin(fabs(xt), fabs(alc - q0)), xt);
  const double br = rsign(dmin(fabs(xt), fabs(alr - q0)), xt);
  const double b0 = bl + br;
  return pos ? q0 + (1.0 - c) * (br - c * b0) : q0 + (1.0 + c) * (bl + c * b0);
}

// Qs: values (shared plane).  XDIR: sweep along i (stride 1) on lines j in [l0, l0+nl), else along j (stride sj) on
// lines i in [l0, l0+nl).  cg: Courant numbers, dxg: cell widths (global planes, same offsets).  Interface values are
// produced for the faces [f0, f1] of every line (the whole line: e.start .. e.end + 1; a strip sweeping along j passes
// its own face rows).  [v0, v1): indices along the sweep direction that may be READ from Qs (resident rows for a y
// sweep; the padded row for an x sweep) — window loads outside it are skipped, their faces are never stored.
// fin(p, value): what to do with the value at plane offset p; it must not write Qs.
template <int MORD, bool XDIR, class Fin>
FV_DEV void ppm_sweep(const Block &b, const double *Qs, int sj, const double *cg, const double *dxg, const Edge1D &e,
                      int l0, int nl, int f0, int f1, int v0, int v1, Fin fin) {
  constexpr int R = SWEEP_R;
  const int st = XDIR ? 1 : sj, ls = XDIR ? sj : 1;
  if (nl <= 0 || f1 < f0) return;  // uniform over the block
  // faces the edge tasks own (skipped by the bulk tasks); empty ranges away from tile edges
  const int elo0 = e.start, elo1 = e.lo ? e.start + 2 : e.start - 1;  // [elo0, elo1]
  const int ehi0 = e.hi ? e.end - 1 : e.end + 2, ehi1 = e.end + 1;    // [ehi0, ehi1]
  // x sweeps: groups start at multiples of R so that the 128-bit window loads are aligned
  const int fb = XDIR ? (f0 & ~(R - 1)) : f0;
  const int ng = (f1 - fb) / R + 1;
  const int nbulk = ng * nl, nedge = (e.lo || e.hi) ? 6 * nl : 0;
  const float inv = 1.0f / (float)(XDIR ? ng : nl);
  b.par(nbulk + nedge, [&](int t) {
    if (t < nbulk) {
      // task -> (line, group): groups fastest for x sweeps (a warp reads whole row segments), lines fastest for y
      // sweeps (consecutive lanes own consecutive columns)
      const int hi_ = (int)(((float)t + 0.5f) * inv);
      const int lo_ = t - hi_ * (XDIR ? ng : nl);
      const int l = l0 + (XDIR ? hi_ : lo_), gi = XDIR ? lo_ : hi_;
      const int F0 = fb + gi * R;
      const int p0 = F0 * st + l * ls;
      // window w[n] = q[F0 - 4 + n], n = 0..9 (w[0] only completes the aligned pair of an x sweep)
      double w[R + 6];
      if (XDIR) {
#pragma unroll
        for (int n = 0; n < R + 6; n += 2) {
          w[n] = w[n + 1] = 0.0;
          if (F0 - 4 + n >= v0 && F0 - 4 + n + 1 < v1) ld_pair(Qs + p0 - 4 + n, w[n], w[n + 1]);
        }
      } else {
#pragma unroll
        for (int n = 1; n < R + 6; ++n) {
          w[n] = 0.0;
          if (F0 - 4 + n >= v0 && F0 - 4 + n < v1) w[n] = Qs[p0 + (n - 4) * st];
        }
      }
      double c[R];
#pragma unroll
      for (int n = 0; n < R; ++n) {
        const int f = F0 + n;
        c[n] = (f >= f0 && f <= f1) ? FV_LDG(cg + p0 + n * st) : 0.0;
      }
      if (MORD < 8) {
        // al at faces F0-1 .. F0+R: al[m] is the edge value at face F0 - 1 + m; q[f] = w[f - F0 + 4]
        double al[R + 2];
#pragma unroll
        for (int m = 0; m < R + 2; ++m) al[m] = PPM_P1 * (w[m + 2] + w[m + 3]) + PPM_P2 * (w[m + 1] + w[m + 4]);
#pragma unroll
        for (int n = 0; n < R; ++n) {
          const int f = F0 + n;
          if (f < f0 || f > f1 || (f >= elo0 && f <= elo1) || (f >= ehi0 && f <= ehi1)) continue;
          fin(p0 + n * st, ppm_face_lt8<MORD>(c[n], w[n + 3], w[n + 4], al[n], al[n + 1], al[n + 2]));
        }
      } else {
        // dm of cells F0-2 .. F0+R: dm[m] belongs to cell F0 - 2 + m = w[m + 2]
        double dm[R + 3];
#pragma unroll
        for (int m = 0; m < R + 3; ++m) dm[m] = ppm_dm8v(w[m + 1], w[m + 2], w[m + 3]);
#pragma unroll
        for (int n = 0; n < R; ++n) {
          const int f = F0 + n;
          if (f < f0 || f > f1 || (f >= elo0 && f <= elo1) || (f >= ehi0 && f <= ehi1)) continue;
          const bool pos = c[n] > 0.0;
          // upwind cell: f - 1 (w[n + 3], dm[n + 1]) for c > 0, else f (w[n + 4], dm[n + 2])
          const double qm = pos ? w[n + 2] : w[n + 3], q0 = pos ? w[n + 3] : w[n + 4], qq = pos ? w[n + 4] : w[n + 5];
          const double dmm = pos ? dm[n] : dm[n + 1], dm0 = pos ? dm[n + 1] : dm[n + 2], dmp = pos ? dm[n + 2] : dm[n + 3];
          fin(p0 + n * st, ppm_face_8(c[n], pos, qm, q0, qq, dmm, dm0, dmp));
        }
      }
    } else {
      // faces next to a cube-tile edge: one-sided edge values / bl, br (xppm.py:148-181, 185-246)
      const int t2 = t - nbulk;
      const int l = l0 + t2 / 6, r = t2 % 6;
      if (r < 3 ? !e.lo : !e.hi) return;
      const int f = r < 3 ? e.start + r : e.end - 1 + (r - 3);
      if (r >= 3 && e.lo && f <= e.start + 2) return;  // tiny domains: already done by the low-edge tasks
      if (f < f0 || f > f1) return;
      const int p = f * st + l * ls;
      auto q = [&](int ii) { return Qs[ii * st + l * ls]; };
      auto dx = [&](int ii) { return dxg[ii * st + l * ls]; };
      if (MORD < 8) {
        auto al = [&](int ii) { return ppm_al_lt8(q, dx, ii, e); };
        fin(p, ppm_flux_staged<MORD>(q, al, dx, cg[p], f, e));
      } else {
        auto dm = [&](int ii) { return ppm_dm8(q, ii); };
        fin(p, ppm_flux_staged<8>(q, dm, dx, cg[p], f, e));
      }
    }
  });
}

}  // namespace fv3
