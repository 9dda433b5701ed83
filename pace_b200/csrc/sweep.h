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

#ifndef FV3_SWEEP_R
#define FV3_SWEEP_R 4
#endif
constexpr int SWEEP_R = FV3_SWEEP_R;  // faces per task (even: the x-sweep windows are read as aligned pairs)

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
// hord 8: q0, dm0 = upwind cell and its limited slope, alc / alr = edge values at its left / right face (xppm.py:82-102,
// 74-79); an edge value is shared by the two cells it separates, so the sweep forms it once per face (measured: tracer
// sub-cycle 3405 -> 3330 us, bit-identical)
FV_HD double ppm_face_8al(double c, bool pos, double q0, double dm0, double alc, double alr) {
  const double xt = 2.0 * dm0;
  const double bl = -1.0 * rsign(dmin(fabs(xt), fabs(alc - q0)), xt);
  const double br = rsign(dmin(fabs(xt), fabs(alr - q0)), xt);
  const double b0 = bl + br;
  return pos ? q0 + (1.0 - c) * (br - c * b0) : q0 + (1.0 + c) * (bl + c * b0);
}

// Generic evaluation of ONE face next to a tile edge (any hord): out of line, one copy per kernel — it is executed by
// a few lanes of tile-edge CTAs only and its one-sided formulas (divides, min / max clamps) are long.
template <int MORD>
#ifndef FV3_HOSTSIM
__device__ __noinline__
#else
inline
#endif
    double
    ppm_edge_face(const double *ql, const double *dl, int st, double c, int f, Edge1D e) {
  auto q = [&](int ii) { return ql[ii * st]; };
  auto dx = [&](int ii) { return dl[ii * st]; };
  if (MORD < 8) {
    auto al = [&](int ii) { return ppm_al_lt8(q, dx, ii, e); };
    return ppm_flux_staged<MORD>(q, al, dx, c, f, e);
  }
  auto dm = [&](int ii) { return ppm_dm8(q, ii); };
  return ppm_flux_staged<8>(q, dm, dx, c, f, e);
}

// Qs: values (shared plane).  XDIR: sweep along i (stride 1) on lines j in [l0, l0+nl), else along j (stride sj) on
// lines i in [l0, l0+nl).  cg: Courant numbers, dxg: cell widths (global planes, same offsets).  Interface values are
// produced for the faces [f0, f1] of every line, e.start <= f0, f1 <= e.end + 1 (the whole line; a strip sweeping along j
// passes its own face rows).
// fin(p, value): what to do with the value at plane offset p; it must not write Qs.
template <int MORD, bool XDIR, class Fin>
struct Sweep {
  const double *Qs, *cg, *dxg;
  // lines outside [alt_lo, alt_hi] read their values from Qalt instead of Qs (same offsets): the y sweeps of the x-halo
  // columns of a strip that holds a cube-corner block (transport.h).  Default: no such lines.
  const double *Qalt;
  int alt_lo, alt_hi;
  Edge1D e;
  int sj, l0, nl, f0, f1, fv0, fb, ng, nbulk, n;  // n: tasks of this sweep (bulk + tile-edge)
  unsigned nfv;
  float inv;
  Fin fin;

  FV_DEV Sweep(const double *Qs_, int sj_, const double *cg_, const double *dxg_, const Edge1D &e_, int l0_, int nl_, int f0_,
               int f1_, Fin fin_)
      : Qs(Qs_), cg(cg_), dxg(dxg_), Qalt(Qs_), alt_lo(-(1 << 30)), alt_hi(1 << 30), e(e_), sj(sj_), l0(l0_), nl(nl_), f0(f0_),
        f1(f1_), fin(fin_) {
    constexpr int R = SWEEP_R;
    // faces the edge tasks own (skipped by the bulk tasks); empty ranges away from tile edges
    const int elo1 = e.lo ? e.start + 2 : e.start - 1;  // [e.start, elo1]
    const int ehi0 = e.hi ? e.end - 1 : e.end + 2;      // [ehi0, e.end + 1]
    // faces the bulk tasks store: [f0, f1] minus the edge faces (contiguous with the ends of the line)
    fv0 = (e.lo && f0 <= elo1) ? elo1 + 1 : f0;
    const int fv1 = (e.hi && f1 >= ehi0) ? ehi0 - 1 : f1;
    const bool any = nl > 0 && f1 >= f0;
    nfv = fv1 >= fv0 ? (unsigned)(fv1 - fv0) : 0u;
    // x sweeps: groups start at multiples of R so that the 128-bit window loads are aligned
    fb = XDIR ? ((R & (R - 1)) == 0 ? (fv0 & ~(R - 1)) : (fv0 & ~1)) : fv0;
    ng = (any && fv1 >= fv0) ? (fv1 - fb) / R + 1 : 0;
    nbulk = ng * nl;
    n = any ? nbulk + ((e.lo || e.hi) ? 2 * nl : 0) : 0;
    inv = 1.0f / (float)(XDIR ? (ng > 0 ? ng : 1) : (nl > 0 ? nl : 1));
  }

  FV_DEV void set_alt(const double *q, int lo, int hi) {
    Qalt = q;
    alt_lo = lo;
    alt_hi = hi;
  }

  FV_DEV void run(int t) const {
    constexpr int R = SWEEP_R;
    const int st = XDIR ? 1 : sj, ls = XDIR ? sj : 1;
    if (t < nbulk) {
      // task -> (line, group): groups fastest for x sweeps (a warp reads whole row segments), lines fastest for y
      // sweeps (consecutive lanes own consecutive columns)
      const int hi_ = (int)(((float)t + 0.5f) * inv);
      const int lo_ = t - hi_ * (XDIR ? ng : nl);
      const int l = l0 + (XDIR ? hi_ : lo_), gi = XDIR ? lo_ : hi_;
      const int F0 = fb + gi * R;
      const int p0 = F0 * st + l * ls;
      const double *Qs = (l < alt_lo || l > alt_hi) ? Qalt : this->Qs;
      // window w[n] = q[F0 - 4 + n], n = 0..9 (w[0] only completes the aligned pair of an x sweep).  Loads are
      // unconditional: a window may reach into the guard doubles around the planes (plane.h) or a neighbouring row,
      // the faces computed from such values lie outside [fv0, fv1] and are dropped.
      double w[R + 6];
      double c[R];
      if (XDIR) {
#pragma unroll
        for (int n = 0; n < R + 6; n += 2) ld_pair(Qs + p0 - 4 + n, w[n], w[n + 1]);
      } else {
#pragma unroll
        for (int n = 1; n < R + 6; ++n) w[n] = Qs[p0 + (n - 4) * st];
      }
#pragma unroll
      for (int n = 0; n < R; ++n) c[n] = FV_LDG(cg + p0 + n * st);
      if (MORD < 8) {
        // al at faces F0-1 .. F0+R: al[m] is the edge value at face F0 - 1 + m; q[f] = w[f - F0 + 4]
        double al[R + 2];
#pragma unroll
        for (int m = 0; m < R + 2; ++m) al[m] = PPM_P1 * (w[m + 2] + w[m + 3]) + PPM_P2 * (w[m + 1] + w[m + 4]);
#pragma unroll
        for (int n = 0; n < R; ++n) {
          if ((unsigned)(F0 + n - fv0) > nfv) continue;
          fin(p0 + n * st, ppm_face_lt8<MORD>(c[n], w[n + 3], w[n + 4], al[n], al[n + 1], al[n + 2]));
        }
      } else {
        // dm of cells F0-2 .. F0+R: dm[m] belongs to cell F0 - 2 + m = w[m + 2]
        double dm[R + 3];
#pragma unroll
        for (int m = 0; m < R + 3; ++m) dm[m] = ppm_dm8v(w[m + 1], w[m + 2], w[m + 3]);
        // edge values at the faces F0-1 .. F0+R, ONCE per face: al[m] (face F0 - 1 + m) is the right edge of cell
        // F0 - 2 + m and the left edge of cell F0 - 1 + m (alc of one cell and alr of its neighbour are the same expression there)
        double al[R + 2];
#pragma unroll
        for (int m = 0; m < R + 2; ++m) al[m] = 0.5 * (w[m + 2] + w[m + 3]) + 1.0 / 3.0 * (dm[m] - dm[m + 1]);
#pragma unroll
        for (int n = 0; n < R; ++n) {
          if ((unsigned)(F0 + n - fv0) > nfv) continue;
          const bool pos = c[n] > 0.0;
          // upwind cell: f - 1 (w[n + 3], dm[n + 1], faces al[n], al[n + 1]) for c > 0, else f (w[n + 4], dm[n + 2], al[n + 1], al[n + 2])
          const double q0 = pos ? w[n + 3] : w[n + 4], dm0 = pos ? dm[n + 1] : dm[n + 2];
          const double alc = pos ? al[n] : al[n + 1], alr = pos ? al[n + 1] : al[n + 2];
          fin(p0 + n * st, ppm_face_8al(c[n], pos, q0, dm0, alc, alr));
        }
      }
    } else {
      // faces next to a cube-tile edge: one-sided edge values / bl, br (xppm.py:148-181, 185-246).  One task per
      // (line, tile edge) = 3 faces.
      const int t2 = t - nbulk;
      const int l = l0 + (t2 >> 1);
      const double *Qs = (l < alt_lo || l > alt_hi) ? Qalt : this->Qs;
      const bool high = t2 & 1;
      if (high ? !e.hi : !e.lo) return;
      const int fa = high ? e.end - 1 : e.start;  // faces fa .. fa + 2
      if (MORD < 8 && fa >= f0 && fa + 2 <= f1 && e.end - e.start >= 6) {
        // straight-line form of compute_al (xppm.py:148-181) around the edge: a = first of the three special faces
        const int a = high ? e.end : e.start - 1;
        const int pa = a * st + l * ls;
        const double *qa = Qs + pa, *da = dxg + pa;
        const double qm4 = high ? qa[-4 * st] : 0.0, qm3 = high ? qa[-3 * st] : 0.0;
        const double qm2 = qa[-2 * st], qm1 = qa[-st], q0 = qa[0], q1 = qa[st], q2 = qa[2 * st], q3 = qa[3 * st];
        const double q4 = high ? 0.0 : qa[4 * st], q5 = high ? 0.0 : qa[5 * st];
        const double dm1 = FV_LDG(da - st), d0 = FV_LDG(da), d1 = FV_LDG(da + st), d2 = FV_LDG(da + 2 * st);
        const double al_a = PPM_C1 * qm2 + PPM_C2 * qm1 + PPM_C3 * q0;
        const double al_a1 = 0.5 * (((2.0 * d0 + dm1) * q0 - d0 * qm1) / (dm1 + d0) + ((2.0 * d1 + d2) * q1 - d1 * q2) / (d1 + d2));
        const double al_a2 = PPM_C3 * q1 + PPM_C2 * q2 + PPM_C1 * q3;
        if (!high) {
          const double al_a3 = PPM_P1 * (q2 + q3) + PPM_P2 * (q1 + q4), al_a4 = PPM_P1 * (q3 + q4) + PPM_P2 * (q2 + q5);
          // faces a+1, a+2, a+3
          fin(pa + st, ppm_face_lt8<MORD>(cg[pa + st], q0, q1, al_a, al_a1, al_a2));
          fin(pa + 2 * st, ppm_face_lt8<MORD>(cg[pa + 2 * st], q1, q2, al_a1, al_a2, al_a3));
          fin(pa + 3 * st, ppm_face_lt8<MORD>(cg[pa + 3 * st], q2, q3, al_a2, al_a3, al_a4));
        } else {
          const double al_m2 = PPM_P1 * (qm3 + qm2) + PPM_P2 * (qm4 + qm1), al_m1 = PPM_P1 * (qm2 + qm1) + PPM_P2 * (qm3 + q0);
          // faces a-1, a, a+1
          fin(pa - st, ppm_face_lt8<MORD>(cg[pa - st], qm2, qm1, al_m2, al_m1, al_a));
          fin(pa, ppm_face_lt8<MORD>(cg[pa], qm1, q0, al_m1, al_a, al_a1));
          fin(pa + st, ppm_face_lt8<MORD>(cg[pa + st], q0, q1, al_a, al_a1, al_a2));
        }
        return;
      }
      for (int r = 0; r < 3; ++r) {
        const int f = fa + r;
        if (high && e.lo && f <= e.start + 2) continue;  // tiny domains: already done by the low-edge task
        if (f < f0 || f > f1) continue;
        const int p = f * st + l * ls;
        fin(p, ppm_edge_face<MORD>(Qs + l * ls, dxg + l * ls, st, cg[p], f, e));
      }
    }
  }
};

template <int MORD, bool XDIR, class Fin>
FV_DEV Sweep<MORD, XDIR, Fin> make_sweep(const double *Qs, int sj, const double *cg, const double *dxg, const Edge1D &e, int l0,
                                        int nl, int f0, int f1, Fin fin) {
  return Sweep<MORD, XDIR, Fin>(Qs, sj, cg, dxg, e, l0, nl, f0, f1, fin);
}

// one sweep = one block-wide pass
template <int MORD, bool XDIR, class Fin>
FV_DEV void ppm_sweep(const Block &b, const double *Qs, int sj, const double *cg, const double *dxg, const Edge1D &e,
                      int l0, int nl, int f0, int f1, Fin fin) {
  const auto sw = make_sweep<MORD, XDIR>(Qs, sj, cg, dxg, e, l0, nl, f0, f1, fin);
  b.par(sw.n, [&](int t) { sw.run(t); });
}
// two independent sweeps (neither reads what the other writes) in ONE block-wide pass: one barrier instead of two and
// twice the tasks to hide latencies behind
template <class S1, class S2>
FV_DEV void ppm_sweep_pair(const Block &b, const S1 &s1, const S2 &s2) {
  b.par(s1.n + s2.n, [&](int t) {
    if (t < s1.n)
      s1.run(t);
    else
      s2.run(t - s1.n);
  });
}

}  // namespace fv3
