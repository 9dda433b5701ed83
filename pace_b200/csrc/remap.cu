// Lagrangian-to-Eulerian vertical remapping: one thread per column, profiles in thread-local arrays.
//   fv3_remap_prep      <- init_pe + moist_cv_pt_pressure + pn2_pk_delp (remapping.py:34-193)
//   fv3_map_single      <- MapSingle.__call__ (map_single.py:147-200): RemapProfile (remap_profile.py:150-563,
//                          kord 9) + lagrangian_contributions (map_single.py:21-81)
//   fv3_fillz           <- FillNegativeTracerValues.__call__ (fillz.py:15-163)
//   fv3_remap_post      <- undo_delz_adjust_and_copy_peln + moist_pkz (remapping.py:46-59, moist_cv.py:112-141)
//   fv3_remap_pressures <- pressures_mapu / pressures_mapv (remapping.py:196-254)
//   fv3_remap_finish    <- update_ua + copy_from_below + moist_pt_last_step / adjust_divide (remapping.py:257-272,
//                          674-695)
//   fv3_fv_setup        <- fv_setup + pt_to_potential_density_pt (moist_cv.py:175-234, fv_dynamics.py:39-52)
#include "common.h"

namespace {

constexpr int NKMAX = 96;
constexpr double RDGAS = 287.05, GRAV = 9.80665, CP_AIR = 1004.6, RVGAS = 461.50;
constexpr double RDG = -RDGAS / GRAV, CV_AIR = CP_AIR - RDGAS, CV_VAP = 3.0 * RVGAS, C_LIQ = 4.1855e3, C_ICE = 1972.0;
constexpr double ZVIR = RVGAS / RDGAS - 1;

struct Profile {
  double a1[NKMAX], a2[NKMAX], a3[NKMAX], a4[NKMAX];
};

FV_HD void posdef_iv1(double &a1, double &a2, double &a3, double &a4) {
  const double da1 = a3 - a2, da2 = da1 * da1, a6da = a4 * da1;
  if (((a1 - a2) * (a1 - a3)) >= 0.0) {
    a2 = a1;
    a3 = a1;
    a4 = 0.0;
  } else if (a6da < -1.0 * da2) {
    a4 = 3.0 * (a2 - a1);
    a3 = a2 - a4;
  } else if (a6da > da2) {
    a4 = 3.0 * (a3 - a1);
    a2 = a3 - a4;
  }
}
FV_HD void remap_constraint(double &a1, double &a2, double &a3, double &a4, bool extm) {
  const double da1 = a3 - a2, da2 = da1 * da1, a6da = a4 * da1;
  if (extm) {
    a2 = a1;
    a3 = a1;
    a4 = 0.0;
  } else if (a6da < -da2) {
    a4 = 3.0 * (a2 - a1);
    a3 = a2 - a4;
  } else if (a6da > da2) {
    a4 = 3.0 * (a3 - a1);
    a2 = a3 - a4;
  }
}
FV_HD void posdef_iv0(double &a1, double &a2, double &a3, double &a4) {
  if (a1 <= 0.0) {
    a2 = a1;
    a3 = a1;
    a4 = 0.0;
  } else if (fabs(a3 - a2) < -a4 && (a1 + 0.25 * ((a3 - a2) * (a3 - a2)) / a4 + a4 * (1.0 / 12.0)) < 0.0) {
    if ((a1 < a3) && (a1 < a2)) {
      a3 = a1;
      a2 = a1;
      a4 = 0.0;
    } else if (a3 > a2) {
      a4 = 3.0 * (a2 - a1);
      a3 = a2 - a4;
    } else {
      a4 = 3.0 * (a3 - a1);
      a2 = a3 - a4;
    }
  }
}
FV_HD double min3(double a, double b, double c) { return (a < b && a < c) ? a : (b < c ? b : c); }
FV_HD double max3(double a, double b, double c) { return (a > b && a > c) ? a : (b > c ? b : c); }

// RemapProfile.__call__ for kord == 9 (remap_profile.py:622-681).  p.a1 holds the layer means on entry.
FV_HD void remap_profile(Profile &p, const double *delp, int km, int iv, double qs, double qmin) {
  double q[NKMAX], gam[NKMAX];
  bool extm[NKMAX];
  const double *a1 = p.a1;
  // set_initial_vals (:150-250)
  if (iv != -2) {
    {
      const double gr = delp[1] / delp[0], bet = gr * (gr + 0.5);
      q[0] = ((gr + gr) * (gr + 1.0) * a1[0] + a1[1]) / bet;
      gam[0] = (1.0 + gr * (gr + 1.5)) / bet;
    }
    for (int k = 1; k < km; ++k) {
      const double d4 = delp[k - 1] / delp[k], bet = 2.0 + d4 + d4 - gam[k - 1];
      q[k] = (3.0 * (a1[k - 1] + d4 * a1[k]) - q[k - 1]) / bet;
      gam[k] = d4 / bet;
    }
    {
      const double d4 = delp[km - 2] / delp[km - 1], a_bot = 1.0 + d4 * (d4 + 1.5);
      q[km] = (2.0 * d4 * (d4 + 1.0) * a1[km - 1] + a1[km - 2] - a_bot * q[km - 1]) / (d4 * (d4 + 0.5) - a_bot * gam[km - 1]);
    }
    for (int k = km - 1; k >= 0; --k) q[k] = q[k] - gam[k] * q[k + 1];
  } else {
    double gr[NKMAX];
    q[0] = 1.5 * a1[0];
    gam[1] = 0.5;
    gr[1] = delp[0] / delp[1];
    q[1] = (3.0 * (a1[0] + a1[1]) - q[0]) / (2.0 + gr[1] + gr[1] - gam[1]);
    for (int k = 2; k < km; ++k) {
      const double old_gr = delp[k - 2] / delp[k - 1], old_bet = 2.0 + old_gr + old_gr - gam[k - 1];
      gam[k] = old_gr / old_bet;
      gr[k] = delp[k - 1] / delp[k];
    }
    for (int k = 2; k < km - 1; ++k) {
      const double bet = 2.0 + gr[k] + gr[k] - gam[k];
      q[k] = (3.0 * (a1[k - 1] + a1[k]) - q[k - 1]) / bet;
    }
    q[km - 1] = (3.0 * (a1[km - 2] + a1[km - 1]) - gr[km - 1] * qs - q[km - 2]) / (2.0 + gr[km - 1] + gr[km - 1] - gam[km - 1]);
    q[km] = qs;
    for (int k = km - 2; k >= 0; --k) q[k] = q[k] - gam[k + 1] * q[k + 1];
  }
  // apply_constraints (:253-337)
  for (int k = 1; k < km; ++k) gam[k] = a1[k] - a1[k - 1];
  for (int k = 1; k < km; ++k) {
    const double tmp = a1[k - 1] > a1[k] ? a1[k - 1] : a1[k], tmp2 = a1[k - 1] < a1[k] ? a1[k - 1] : a1[k];
    if (k == 1 || k == km - 1) {
      if (q[k] >= tmp) q[k] = tmp;
      if (q[k] <= tmp2) q[k] = tmp2;
    } else if (gam[k - 1] * gam[k + 1] > 0) {
      if (q[k] >= tmp) q[k] = tmp;
      if (q[k] <= tmp2) q[k] = tmp2;
    } else if (gam[k - 1] > 0) {
      if (q[k] <= tmp2) q[k] = tmp2;
    } else {
      if (q[k] >= tmp) q[k] = tmp;
      if (iv == 0 && q[k] < 0.0) q[k] = 0.0;
    }
  }
  for (int k = 0; k < km; ++k) {
    p.a2[k] = q[k];
    p.a3[k] = q[k + 1];
  }
  extm[0] = (p.a2[0] - a1[0]) * (p.a3[0] - a1[0]) > 0.0;
  for (int k = 1; k < km - 1; ++k) extm[k] = gam[k] * gam[k + 1] < 0.0;
  extm[km - 1] = (p.a2[km - 1] - a1[km - 1]) * (p.a3[km - 1] - a1[km - 1]) > 0.0;
  // set_interpolation_coefficients (:340-563)
  if (iv == 0 && p.a2[0] < 0.0) p.a2[0] = 0.0;
  if (iv == -1 && p.a2[0] * a1[0] <= 0.0) p.a2[0] = 0.0;
  for (int k = 0; k < 2; ++k) p.a4[k] = 3.0 * (2.0 * a1[k] - (p.a2[k] + p.a3[k]));
  posdef_iv1(p.a1[0], p.a2[0], p.a3[0], p.a4[0]);
  remap_constraint(p.a1[1], p.a2[1], p.a3[1], p.a4[1], extm[1]);
  for (int k = 2; k < km - 2; ++k) {
    const double v1 = a1[k];
    const double pmp_1 = v1 - 2.0 * gam[k + 1], lac_1 = pmp_1 + 1.5 * gam[k + 2];
    const double pmp_2 = v1 + 2.0 * gam[k], lac_2 = pmp_2 - 1.5 * gam[k - 1];
    if ((extm[k] && extm[k - 1]) || (extm[k] && extm[k + 1]) || (extm[k] && (qmin > 0.0 && v1 < qmin))) {
      p.a2[k] = v1;
      p.a3[k] = v1;
      p.a4[k] = 0.0;
    } else {
      p.a4[k] = 6.0 * v1 - 3.0 * (p.a2[k] + p.a3[k]);
      if (fabs(p.a4[k]) > fabs(p.a2[k] - p.a3[k])) {
        double tmin = min3(v1, pmp_1, lac_1), tmax = max3(v1, pmp_1, lac_1);
        double t0 = p.a2[k] > tmin ? p.a2[k] : tmin;
        p.a2[k] = t0 < tmax ? t0 : tmax;
        tmin = min3(v1, pmp_2, lac_2);
        tmax = max3(v1, pmp_2, lac_2);
        t0 = p.a3[k] > tmin ? p.a3[k] : tmin;
        p.a3[k] = t0 < tmax ? t0 : tmax;
        p.a4[k] = 6.0 * v1 - 3.0 * (p.a2[k] + p.a3[k]);
      }
    }
    if (iv == 0) posdef_iv0(p.a1[k], p.a2[k], p.a3[k], p.a4[k]);
  }
  if (iv == 0 && p.a3[km - 1] < 0.0) p.a3[km - 1] = 0.0;
  if (iv == -1 && p.a3[km - 1] * a1[km - 1] <= 0.0) p.a3[km - 1] = 0.0;
  for (int k = km - 2; k < km; ++k) p.a4[k] = 3.0 * (2.0 * a1[k] - (p.a2[k] + p.a3[k]));
  remap_constraint(p.a1[km - 2], p.a2[km - 2], p.a3[km - 2], p.a4[km - 2], extm[km - 2]);
  posdef_iv1(p.a1[km - 1], p.a2[km - 1], p.a3[km - 1], p.a4[km - 1]);
}

FV_HD void moist_cv(double qv, double ql_, double qr, double qs_, double qi, double qg, double &cvm, double &gz) {
  const double ql = ql_ + qr, qs = qi + qs_ + qg;
  gz = ql + qs;
  cvm = (1.0 - (qv + gz)) * CV_AIR + qv * CV_VAP + ql * C_LIQ + qs * C_ICE;
}

}  // namespace

extern "C" {

int fv3_map_single(fv3_ctx *ctx, double *q1, const double *pe1, const double *pe2, const double *qs, int qs_is_2d,
                   double qmin, int kord, int iv, int i_extra, int j_extra, void *stream) {
  const fv3_geom g = ctx->g;
  if (kord < 0) kord = -kord;
  if (kord != 9) {
    fv3::set_error("fv3_map_single: only kord 9 is implemented");
    return -1;
  }
  if (g.nz + 1 > NKMAX) {
    fv3::set_error("fv3_map_single: nz too large");
    return -1;
  }
  const int h = g.halo, km = g.nz;
  fv3::launch2d(ctx, (cudaStream_t)stream, h, h + g.nx + i_extra, h, h + g.ny + j_extra, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    Profile p;
    double dp1[NKMAX], p1[NKMAX];
    for (int k = 0; k <= km; ++k) p1[k] = pe1[c0 + k * sk];
    for (int k = 0; k < km; ++k) {
      p.a1[k] = q1[c0 + k * sk];
      dp1[k] = p1[k + 1] - p1[k];
    }
    double qsv = 0.0;
    if (qs != nullptr) qsv = qs_is_2d ? qs[O2(s, i, j)] : qs[c0];
    remap_profile(p, dp1, km, iv, qsv, qmin);
    // lagrangian_contributions (map_single.py:21-81); L = absolute source layer
    int L = 0;
    double top = pe2[c0];
    for (int k = 0; k < km; ++k) {
      const double bot = pe2[c0 + (k + 1) * sk];
      const double pl = (top - p1[L]) / dp1[L];
      double out;
      if (bot <= p1[L + 1]) {
        const double pr = (bot - p1[L]) / dp1[L];
        out = p.a2[L] + 0.5 * (p.a4[L] + p.a3[L] - p.a2[L]) * (pr + pl) - p.a4[L] * 1.0 / 3.0 * (pr * (pr + pl) + pl * pl);
      } else {
        double qsum = (p1[L + 1] - top) * (p.a2[L] + 0.5 * (p.a4[L] + p.a3[L] - p.a2[L]) * (1.0 + pl) -
                                           p.a4[L] * 1.0 / 3.0 * (1.0 + pl * (1.0 + pl)));
        L = L + 1;
        while (L + 1 <= km && p1[L + 1] < bot) {
          qsum += dp1[L] * p.a1[L];
          L = L + 1;
        }
        if (L > km - 1) L = km - 1;
        const double dp = bot - p1[L], esl = dp / dp1[L];
        qsum += dp * (p.a2[L] + 0.5 * esl * (p.a3[L] - p.a2[L] + p.a4[L] * (1.0 - (2.0 / 3.0) * esl)));
        out = qsum / (bot - top);
      }
      q1[c0 + k * sk] = out;
      top = bot;
    }
  });
  return fv3::check_launch("fv3_map_single");
}

int fv3_fillz(fv3_ctx *ctx, double *const *tracers, int nq, const double *dp2, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, km = g.nz;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, nq, FV_LAMBDA(int s, int i, int j, int t) { FV_DEV_GM
    double *qf = tracers[t];
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    double q[NKMAX], dp[NKMAX], lower_fix[NKMAX], upper_fix[NKMAX], dm[NKMAX], dm_pos[NKMAX];
    bool any_neg = false;
    for (int k = 0; k < km; ++k) {
      q[k] = qf[c0 + k * sk];
      dp[k] = dp2[c0 + k * sk];
      lower_fix[k] = 0.0;
      upper_fix[k] = 0.0;
      any_neg = any_neg || q[k] < 0.0;
    }
    int zfix = 0;
    if (q[0] < 0.0) q[1] = q[1] + q[0] * dp[0] / dp[1];
    if (q[0] < 0) q[0] = 0;
    dm[0] = q[0] * dp[0];
    for (int k = 1; k < km - 1; ++k) {
      if (lower_fix[k - 1] != 0.0) q[k] = q[k] - (lower_fix[k - 1] / dp[k]);
      if (q[k] < 0.0) {
        zfix += 1;
        if (q[k - 1] > 0.0) {
          const double a = q[k - 1] * dp[k - 1], b = -(q[k] * dp[k]);
          const double dq = a < b ? a : b;
          q[k] = q[k] + dq / dp[k];
          upper_fix[k] = dq;
        }
        if ((q[k] < 0.0) && (q[k + 1] > 0.0)) {
          const double a = q[k + 1] * dp[k + 1], b = -(q[k] * dp[k]);
          const double dq = a < b ? a : b;
          q[k] = q[k] + dq / dp[k];
          lower_fix[k] = dq;
        }
      }
    }
    for (int k = 0; k < km - 1; ++k) {
      if (upper_fix[k + 1] != 0.0) q[k] = q[k] - upper_fix[k + 1] / dp[k];
      dm[k] = q[k] * dp[k];
      dm_pos[k] = dm[k] > 0.0 ? dm[k] : 0.0;
    }
    {
      const int k = km - 1;
      if (lower_fix[k - 1] != 0.0) q[k] = q[k] - (lower_fix[k - 1] / dp[k]);
      const double qup = q[k - 1] * dp[k - 1], qly = -q[k] * dp[k];
      const double dup = qup < qly ? qup : qly;
      if ((q[k] < 0.0) && (q[k - 1] > 0.0)) {
        zfix += 1;
        q[k] = q[k] + (dup / dp[k]);
        upper_fix[k] = dup;
      }
      dm[k] = q[k] * dp[k];
      dm_pos[k] = dm[k] > 0.0 ? dm[k] : 0.0;
    }
    {
      const int k = km - 2;
      if (upper_fix[k + 1] != 0.0) {
        q[k] = q[k] - (upper_fix[k + 1] / dp[k]);
        dm[k] = q[k] * dp[k];
        dm_pos[k] = dm[k] > 0.0 ? dm[k] : 0.0;
      }
    }
    double sum0 = 0.0, sum1 = 0.0;
    for (int k = 1; k < km; ++k) {
      sum0 += dm[k];
      sum1 += dm_pos[k];
    }
    const double fac = sum0 > 0.0 ? sum0 / sum1 : 0.0;
    if (zfix > 0 && fac > 0.0)
      for (int k = 1; k < km; ++k) {
        const double v = fac * dm[k] / dp[k];
        q[k] = v > 0.0 ? v : 0.0;
      }
    (void)any_neg;
    for (int k = 0; k < km; ++k) qf[c0 + k * sk] = q[k];
  });
  return fv3::check_launch("fv3_fillz");
}

// tracers6: qvapor, qliquid, qrain, qsnow, qice, qgraupel (device array of 6 pointers)
int fv3_remap_prep(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *pt, double *cappa, double *delp,
                   double *delz, const double *pe, double *pe1, double *pe2, double *dp2, double *ps, double *pn2,
                   const double *peln, double *pk, double ptop, double akap, double r_vir, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo, km = g.nz;
  fv3::launch2d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny + 1, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    // init_pe on the (nx, ny+1) domain
    for (int k = 0; k <= km; ++k) pe1[c0 + k * sk] = pe[c0 + k * sk];
    pe2[c0] = ptop;
    pe2[c0 + km * sk] = pe[c0 + km * sk];
    if (j >= h + g.ny) return;
    const double psv = pe[c0 + km * sk];
    ps[O2(s, i, j)] = psv;
    for (int k = 0; k < km; ++k) {
      const int64_t o = c0 + k * sk;
      double cvm, gz;
      moist_cv(tracers6[0][o], tracers6[1][o], tracers6[2][o], tracers6[3][o], tracers6[4][o], tracers6[5][o], cvm, gz);
      q_con[o] = gz;
      const double cp = RDGAS / (RDGAS + cvm / (1.0 + r_vir * tracers6[0][o]));
      cappa[o] = cp;
      pt[o] = pt[o] * exp(cp / (1.0 - cp) * log(RDG * delp[o] / delz[o] * pt[o]));
      delz[o] = -delz[o] / delp[o];
    }
    for (int k = 1; k < km; ++k) pe2[c0 + k * sk] = m.ak[k] + m.bk[k] * psv;
    pn2[c0 + km * sk] = peln[c0 + km * sk];
    for (int k = 0; k < km; ++k) {
      const int64_t o = c0 + k * sk;
      const double d = pe2[o + sk] - pe2[o];
      dp2[o] = d;
      delp[o] = d;
      const double l = log(pe2[o]);
      pn2[o] = l;
      pk[o] = exp(akap * l);
    }
  });
  return fv3::check_launch("fv3_remap_prep");
}

int fv3_remap_post(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *pkz, const double *pt, double *cappa,
                   const double *delp, double *delz, double *peln, double *pe0, const double *pn2, double r_vir,
                   void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, km = g.nz;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, km + 1, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    pe0[o] = peln[o];
    peln[o] = pn2[o];
    if (k == km) return;
    const double dz = -delz[o] * delp[o];
    delz[o] = dz;
    double cvm, gz;
    moist_cv(tracers6[0][o], tracers6[1][o], tracers6[2][o], tracers6[3][o], tracers6[4][o], tracers6[5][o], cvm, gz);
    q_con[o] = gz;
    const double cp = RDGAS / (RDGAS + cvm / (1.0 + r_vir * tracers6[0][o]));
    cappa[o] = cp;
    pkz[o] = exp(cp * log(RDG * delp[o] / dz * pt[o]));
  });
  return fv3::check_launch("fv3_remap_post");
}

// dir 0: pressures_mapu on (nx, ny+1); dir 1: pressures_mapv on (nx+1, ny)
int fv3_remap_pressures(fv3_ctx *ctx, const double *pe, double *pe0, double *pe3, int dir, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const int h = g.halo, km = g.nz;
  const int64_t off = dir == 0 ? g.sj : 1;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx + (dir == 1), h, h + g.ny + (dir == 0), 0, km + 1, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k), ob = O3(s, i, j, km);
    const double bkh = 0.5 * m.bk[k];
    if (k == 0) {
      pe0[o] = pe[o];
      pe3[o] = dir == 0 ? m.ak[0] + bkh * (pe[ob - off] + pe[ob]) : m.ak[0];
    } else {
      pe0[o] = 0.5 * (pe[o - off] + pe[o]);
      pe3[o] = m.ak[k] + bkh * (pe[ob - off] + pe[ob]);
    }
  });
  return fv3::check_launch("fv3_remap_pressures");
}

int fv3_remap_finish(fv3_ctx *ctx, double *const *tracers6, const double *pe2, double *pe, double *pt, const double *pkz,
                     int last_step, double dtmp, double r_vir, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo, km = g.nz;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, km, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    if (k >= 1) pe[o] = pe2[o];
    if (last_step) {
      const double gz = tracers6[1][o] + tracers6[2][o] + tracers6[4][o] + tracers6[3][o] + tracers6[5][o];
      pt[o] = (pt[o] + dtmp * pkz[o]) / ((1.0 + r_vir * tracers6[0][o]) * (1.0 - gz));
    } else {
      pt[o] = pt[o] / pkz[o];
    }
  });
  return fv3::check_launch("fv3_remap_finish");
}

// compute_preamble (fv_dynamics.py:440-483): fv_setup + pt_to_potential_density_pt
int fv3_fv_setup(fv3_ctx *ctx, double *const *tracers6, double *q_con, double *cvm_out, double *pkz, double *pt,
                 double *cappa, const double *delp, const double *delz, double *dp1, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    double cvm, qc;
    moist_cv(tracers6[0][o], tracers6[1][o], tracers6[2][o], tracers6[3][o], tracers6[4][o], tracers6[5][o], cvm, qc);
    cvm_out[o] = cvm;
    q_con[o] = qc;
    const double d1 = ZVIR * tracers6[0][o];
    dp1[o] = d1;
    const double cp = RDGAS / (RDGAS + cvm / (1.0 + d1));
    cappa[o] = cp;
    const double pz = exp(cp * log(RDG * delp[o] * pt[o] * (1.0 + d1) * (1.0 - qc) / delz[o]));
    pkz[o] = pz;
    pt[o] = pt[o] * (1.0 + d1) * (1.0 - qc) / pz;
  });
  return fv3::check_launch("fv3_fv_setup");
}

// omega_from_w (fv_dynamics.py:55-64)
int fv3_omega_from_w(fv3_ctx *ctx, const double *delp, const double *delz, const double *w, double *omga, void *stream) {
  const fv3_geom g = ctx->g;
  const int h = g.halo;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, 0, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k);
    omga[o] = delp[o] / delz[o] * w[o];
  });
  return fv3::check_launch("fv3_omega_from_w");
}

}  // extern "C"
