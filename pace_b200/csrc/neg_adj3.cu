// Negative mixing-ratio adjustment at the end of the timestep.
//   fv3_neg_adj3 <- AdjustNegativeTracerMixingRatio.__call__ (fv3core/pace/fv3core/stencils/neg_adj3.py:377-420):
//                   fix_neg_water (:98-143, with fix_negative_ice :15-54 and fix_negative_liq :57-95), fillq on
//                   qgraupel and qrain (:146-175), fix_water_vapor_down (:179-249), fix_neg_cloud (:252-278)
// One thread per column, the five stencils back to back in the reference's order, statement by statement (GT4Py
// semantics: a FORWARD / BACKWARD computation finishes a level before the next, a PARALLEL one reads only values of
// earlier computations).  fillq, fix_water_vapor_down and fix_neg_cloud leave a column without negative values
// bit-identical (checked against the reference by oracle/refshim/gen_neg_adj.py), so such columns skip their sweeps.
#include "common.h"

namespace {

constexpr int NKMAX = 96;
constexpr double RDGAS = 287.05, RVGAS = 461.50, CP_AIR = 1004.6, HLV = 2.5e6, HLF = 3.3358e5, TICE = 273.16;
constexpr double CV_AIR = CP_AIR - RDGAS, CV_VAP = 3.0 * RVGAS, C_ICE = 1972.0, C_LIQ = 4.1855e3;
constexpr double DC_ICE = C_LIQ - C_ICE, LI0 = HLF - DC_ICE * TICE;
constexpr double D0_VAP = CV_VAP - C_LIQ, LV00 = HLV - D0_VAP * TICE;  // non-hydrostatic (:352-353)

struct Water {
  double qvapor, qice, qsnow, qgraupel, qrain, qliquid, pt;
};

FV_HD void fix_negative_ice(Water &w, double lcpk, double icpk) {
  double dq;
  const double qsum = w.qice + w.qsnow;
  if (qsum > 0.0) {
    if (w.qice < 0.0) {
      w.qice = 0.0;
      w.qsnow = qsum;
    } else if (w.qsnow < 0.0) {
      w.qsnow = 0.0;
      w.qice = qsum;
    }
  } else {
    w.qice = 0.0;
    w.qsnow = 0.0;
    w.qgraupel = w.qgraupel + qsum;
  }
  if (w.qgraupel < 0.0) {
    dq = w.qsnow < -w.qgraupel ? w.qsnow : -w.qgraupel;
    w.qsnow = w.qsnow - dq;
    w.qgraupel = w.qgraupel + dq;
    if (w.qgraupel < 0.0) {
      dq = w.qice < -w.qgraupel ? w.qice : -w.qgraupel;
      w.qice = w.qice - dq;
      w.qgraupel = w.qgraupel + dq;
    }
  }
  if (w.qgraupel < 0.0 && w.qrain > 0.0) {
    dq = w.qrain < -w.qgraupel ? w.qrain : -w.qgraupel;
    w.qgraupel = w.qgraupel + dq;
    w.qliquid = w.qliquid - dq;
    w.pt = w.pt + dq * icpk;
  }
  if (w.qgraupel < 0.0 && w.qliquid > 0.0) {
    dq = w.qliquid < -w.qgraupel ? w.qliquid : -w.qgraupel;
    w.qgraupel = w.qgraupel + dq;
    w.qliquid = w.qliquid - dq;
    w.pt = w.pt + dq * icpk;
  }
  if (w.qgraupel < 0.0 && w.qvapor > 0.0) {
    dq = 0.999 * w.qvapor < -w.qgraupel ? 0.999 * w.qvapor : -w.qgraupel;
    w.qgraupel = w.qgraupel + dq;
    w.qvapor = w.qvapor - dq;
    w.pt = w.pt + dq * (icpk + lcpk);
  }
}

FV_HD void fix_negative_liq(Water &w, double lcpk, double icpk) {
  double dq;
  const double qsum = w.qliquid + w.qrain;
  const double pos_qgraupel = 0.0 > w.qgraupel ? 0.0 : w.qgraupel;
  double qrain_tmp = 0.0, dq1 = 0.0;
  if (qsum > 0.0) {
    if (w.qrain < 0.0) {
      w.qrain = 0.0;
      w.qliquid = qsum;
    } else if (w.qliquid < 0.0) {
      w.qliquid = 0.0;
      w.qrain = qsum;
    }
  } else {
    w.qliquid = 0.0;
    qrain_tmp = qsum;
    dq = pos_qgraupel < -qrain_tmp ? pos_qgraupel : -qrain_tmp;
    qrain_tmp = qrain_tmp + dq;
    w.qgraupel = w.qgraupel - dq;
    w.pt = w.pt - dq * icpk;
    if (w.qrain < 0.0) {
      dq = (w.qice + w.qsnow) < -qrain_tmp ? w.qice + w.qsnow : -qrain_tmp;
      qrain_tmp = qrain_tmp + dq;
      dq1 = dq < w.qsnow ? dq : w.qsnow;
      w.qsnow = w.qsnow - dq1;
      w.qice = w.qice + dq1 - dq;
      w.pt = w.pt - dq * icpk;
    }
    w.qrain = qrain_tmp;
    if (w.qrain < 0.0 && w.qvapor > 0.0) {
      dq = 0.999 * w.qvapor < -w.qrain ? 0.999 * w.qvapor : -w.qrain;
      w.qvapor = w.qvapor - dq;
      w.qrain = w.qrain + dq;
      w.pt = w.pt + dq * lcpk;
    }
  }
}

// fillq (:146-175): q in place, one column
FV_HD void fillq_column(double *q, const double *dp, int64_t sk, int km) {
  double sum1 = 0.0, sum2 = 0.0;
  bool neg = false;
  for (int k = 0; k < km; ++k) {
    const double qv = q[k * sk];
    if (qv > 0) sum1 = sum1 + qv * dp[k * sk];
    neg = neg || qv < 0.0;
  }
  if (!neg) return;
  for (int k = km - 1; k >= 0; --k) {
    const double qv = q[k * sk], d = dp[k * sk];
    if (qv < 0.0 && sum1 >= 0) {
      const double dq = sum1 < -qv * d ? sum1 : -qv * d;
      sum1 = sum1 - dq;
      sum2 = sum2 + dq;
      q[k * sk] = qv + dq / d;
    }
  }
  for (int k = km - 1; k >= 0; --k) {
    const double qv = q[k * sk], d = dp[k * sk];
    if (qv > 0.0 && sum1 >= 1e-12 && sum2 > 0) {
      const double dq = sum2 < qv * d ? sum2 : qv * d;
      sum2 = sum2 - dq;
      q[k * sk] = qv - dq / d;
    }
  }
}

}  // namespace

extern "C" {

int fv3_neg_adj3(fv3_ctx *ctx, double *qvapor, double *qliquid, double *qrain, double *qsnow, double *qice,
                 double *qgraupel, double *qcld, double *pt, const double *delp, void *stream) {
  const fv3_geom g = ctx->g;
  if (g.nz + 1 > NKMAX || g.nz < 4) {
    fv3::set_error("fv3_neg_adj3: nz out of range");
    return -1;
  }
  const int h = g.halo, km = g.nz;
  fv3::launch2d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, FV_LAMBDA(int s, int i, int j) { FV_DEV_GM
    const int64_t c0 = O3(s, i, j, 0), sk = g.sk;
    // fix_neg_water (:98-143): every level on its own
    bool vneg = false, cneg = false;
    for (int k = 0; k < km; ++k) {
      const int64_t o = c0 + k * sk;
      Water w{qvapor[o], qice[o], qsnow[o], qgraupel[o], qrain[o], qliquid[o], pt[o]};
      cneg = cneg || qcld[o] < 0.0;
      // a level whose six species are all non-negative comes out of fix_negative_ice / _liq unchanged
      if (w.qvapor >= 0.0 && w.qice >= 0.0 && w.qsnow >= 0.0 && w.qgraupel >= 0.0 && w.qrain >= 0.0 && w.qliquid >= 0.0) continue;
      const double q_liq = 0.0 > w.qliquid + w.qrain ? 0.0 : w.qliquid + w.qrain;
      const double q_sol = 0.0 > w.qice + w.qsnow ? 0.0 : w.qice + w.qsnow;
      const double cpm = (1.0 - (w.qvapor + q_liq + q_sol)) * CV_AIR + w.qvapor * CV_VAP + q_liq * C_LIQ + q_sol * C_ICE;
      const double lcpk = (LV00 + D0_VAP * w.pt) / cpm;
      const double icpk = (LI0 + DC_ICE * w.pt) / cpm;
      fix_negative_ice(w, lcpk, icpk);
      fix_negative_liq(w, lcpk, icpk);
      qvapor[o] = w.qvapor;
      qice[o] = w.qice;
      qsnow[o] = w.qsnow;
      qgraupel[o] = w.qgraupel;
      qrain[o] = w.qrain;
      qliquid[o] = w.qliquid;
      pt[o] = w.pt;
      vneg = vneg || w.qvapor < 0.0;
    }
    const double *dp = delp + c0;
    fillq_column(qgraupel + c0, dp, sk, km);
    fillq_column(qrain + c0, dp, sk, km);
    // fix_water_vapor_down (:179-249)
    if (vneg) {
      double *q = qvapor + c0;
      double upper_fix[NKMAX], lower_fix[NKMAX];
      for (int k = 0; k < km; ++k) {
        upper_fix[k] = 0.0;
        lower_fix[k] = 0.0;
      }
      if (q[0] < 0) q[sk] = q[sk] + q[0] * dp[0] / dp[sk];
      if (q[0] < 0.0) q[0] = 0.0;
      for (int k = 1; k < km - 1; ++k) {
        double qk = q[k * sk];
        const double qm = q[(k - 1) * sk], d = dp[k * sk];
        double dq = qm * dp[(k - 1) * sk];
        if (lower_fix[k - 1] != 0) qk += lower_fix[k - 1] / d;
        if ((qk < 0) && (qm > 0)) {
          dq = dq < -qk * d ? dq : -qk * d;
          upper_fix[k] = dq;
          qk += dq / d;
        }
        if (qk < 0) {
          lower_fix[k] = qk * d;
          qk = 0;
        }
        q[k * sk] = qk;
      }
      for (int k = 0; k < km - 2; ++k)
        if (upper_fix[k + 1] != 0) q[k * sk] = q[k * sk] - upper_fix[k + 1] / dp[k * sk];
      const int kb = km - 1;
      if (lower_fix[kb - 1] > 0) q[kb * sk] = q[kb * sk] + lower_fix[kb] / dp[kb * sk];
      // the bottom value is carried up the column (the reference's upper_fix re-use) and filled from the levels above
      double qbot = q[kb * sk];
      const double dp_bottom = dp[kb * sk];
      for (int k = km - 2; k >= 0; --k) {
        const double qk = q[k * sk], d = dp[k * sk];
        double dq = qk * d;
        if ((qbot < 0) && (qk > 0)) {
          if (dq >= -qbot * dp_bottom) dq = -qbot * dp_bottom;
          q[k * sk] = qk - dq / d;
          qbot = qbot + dq / dp_bottom;
        }
      }
      q[kb * sk] = qbot;
    }
    // fix_neg_cloud (:252-278)
    if (cneg) {
      double *q = qcld + c0;
      for (int k = 1; k < km - 1; ++k)
        if (q[(k - 1) * sk] < 0.0) q[k * sk] = q[k * sk] + q[(k - 1) * sk] * dp[(k - 1) * sk] / dp[k * sk];
      for (int k = 1; k < km - 1; ++k)
        if (q[k * sk] < 0.0) q[k * sk] = 0.0;
      {
        const int k = km - 2;
        const double qk = q[k * sk], qn = q[(k + 1) * sk];
        if (qn < 0.0 && qk > 0) {
          const double a = -qk * dp[k * sk], b = qn * dp[(k + 1) * sk];
          const double dq = a < b ? a : b;
          q[k * sk] = qk - dq / dp[k * sk];
        }
      }
      {
        const int k = km - 1;
        double qk = q[k * sk];
        const double qm = q[(k - 1) * sk];
        if (qk < 0 && qm > 0.0) {
          const double a = -qk * dp[k * sk], b = qm * dp[(k - 1) * sk];
          const double dq = a < b ? a : b;
          qk = qk + dq / dp[k * sk];
          qk = 0.0 > qk ? 0.0 : qk;
          q[k * sk] = qk;
        }
      }
    }
  });
  return fv3::check_launch("fv3_neg_adj3");
}

}  // extern "C"
