// Fast saturation adjustment of the vertical remap (GFDL cloud microphysics, Zhou, Harris and Chen 2022).
//   fv3_sat_adjust <- SatAdjust3d.__call__ (fv3core/pace/fv3core/stencils/saturation_adjustment.py:945-1108), the
//                     `satadjust` stencil (:561-943) and its gtscript functions (:30-560), non-hydrostatic branch.
// Pointwise: one thread per cell of levels [kmp, nz), every statement in the reference's order.  The reference evaluates
// its saturation "tables" on the fly (qs_tablew_fn / qs_table2_fn / des2_table / desw_table are closed-form
// functions of the table index, :70-167), and so does this kernel: nothing is tabulated, a lookup is a few exp / log of
// the index temperature.
#include "common.h"

namespace {

// util/pace/util/constants.py:30-97 (GFS_PHYS = True)
constexpr double GRAV = 9.80665, RDGAS = 287.05, RVGAS = 461.50, HLV = 2.5e6, HLF = 3.3358e5, CP_AIR = 1004.6;
constexpr double CV_AIR = CP_AIR - RDGAS, RDG = -RDGAS / GRAV, CV_VAP = 3.0 * RVGAS, C_ICE = 1972.0, C_LIQ = 4.1855e3;
constexpr double CP_VAP = 4.0 * RVGAS, TICE = 273.16, DC_ICE = C_LIQ - C_ICE, DC_VAP = CP_VAP - C_LIQ, D2ICE = DC_VAP + DC_ICE;
constexpr double LV0 = HLV - DC_VAP * TICE, LI00 = HLF - DC_ICE * TICE, LI2 = LV0 + LI00, E00 = 611.21, T_WFR = TICE - 40.0;
constexpr double TICE0 = TICE - 0.01, T_SAT_MIN = TICE - 160.0, LAT2 = (HLV + HLF) * (HLV + HLF);
constexpr double DELT = 0.1;
constexpr int QS_LENGTH = 2621;

FV_HD double dimf(double a, double b) { return a - b > 0 ? a - b : 0.0; }  // basic_operations.dim
FV_HD double tem_lower(double i) { return T_SAT_MIN + DELT * i; }
FV_HD double tem_upper(double i) { return 253.16 + DELT * i; }
FV_HD double q_table_oneline(double dhc, double lhc, double tem) {
  return E00 * exp((dhc * log(tem / TICE) + (tem - TICE) / (tem * TICE) * lhc) / RVGAS);
}
FV_HD double table_vapor(double tem) { return q_table_oneline(DC_VAP, LV0, tem); }
FV_HD double table_ice(double tem) { return q_table_oneline(D2ICE, LI2, tem); }

// qs_table2_fn (:93-127): es over ice below index 1600, over water above, smoothed at 1599 / 1600
FV_HD double qs_table2(double i) {
  double tem0 = tem_lower(i);
  double table2 = i < 1600 ? table_ice(tem0) : table_vapor(tem0);
  if (i == 1599) {
    double table = table_ice(tem0);
    tem0 = tem_upper(i - 1400);
    table = (0.05 * (TICE - tem0)) * table + (0.05 * (tem0 - 253.16)) * table_vapor(tem0);
    const double m1 = table_ice(tem_lower(1598)), p1 = table_vapor(tem_lower(1600));
    table2 = 0.25 * (m1 + 2.0 * table + p1);
  }
  if (i == 1600) {
    const double table = table_vapor(tem_upper(i - 1400));
    const double m1 = table_ice(tem_lower(1599)), p1 = table_vapor(tem_lower(1601));
    table2 = 0.25 * (m1 + 2.0 * table + p1);
  }
  return table2;
}
FV_HD double qs_tablew(double i) { return table_vapor(tem_lower(i)); }
// des2_table / desw_table (:146-167) with des_end (:133-139)
FV_HD double des2_table(double i) {
  const double t = qs_table2(i);
  double des2 = fv3::dmax(0.0, qs_table2(i + 1) - t);
  if (i == QS_LENGTH - 1) des2 = fv3::dmax(0.0, t - qs_table2(i - 1));
  return des2;
}
FV_HD double desw_table(double i) {
  const double t = qs_tablew(i);
  double desw = fv3::dmax(0.0, qs_tablew(i + 1) - t);
  if (i == QS_LENGTH - 1) desw = fv3::dmax(0.0, t - qs_table2(i - 1));  // des_end uses qs_table2_fn for both tables
  return desw;
}
FV_HD double ap1_for_wqs2(double ta) {
  const double ap1 = 10.0 * dimf(ta, T_SAT_MIN) + 1.0;
  return fv3::dmin(ap1, (double)QS_LENGTH) - 1;
}
// wqs2_fn_w / wqs2_fn_2 (:468-493): saturation mixing ratio and its temperature derivative
template <bool WATER>
FV_HD void wqs2(double ta, double den, double &wqsat, double &dqdt) {
  const double ap1 = ap1_for_wqs2(ta);
  const double it = floor(ap1), it2 = floor(ap1 - 0.5), it2_p1 = it2 + 1;
  const double tab = WATER ? qs_tablew(it) : qs_table2(it);
  const double des = WATER ? desw_table(it) : des2_table(it);
  const double des_2 = WATER ? desw_table(it2) : des2_table(it2);
  const double des_p1 = WATER ? desw_table(it2_p1) : des2_table(it2_p1);
  const double es = tab + (ap1 - it) * des;
  const double denom = RVGAS * ta * den;
  wqsat = es / denom;
  dqdt = 10.0 * (des_2 + (ap1 - it2) * (des_p1 - des_2));
  dqdt = dqdt / denom;
}
template <bool WATER>
FV_HD double wqs1(double it, double ap1, double ta, double den) {
  const double tab = WATER ? qs_tablew(it) : qs_table2(it);
  const double des = WATER ? desw_table(it) : des2_table(it);
  const double es = tab + (ap1 - it) * des;
  return es / (RVGAS * ta * den);
}
FV_HD double compute_cvm(double mc_air, double qv, double c_vap, double q_liq, double q_sol) {
  return mc_air + qv * c_vap + q_liq * C_LIQ + q_sol * C_ICE;
}
FV_HD double minmax_tmp_h20(double qa, double qb) { return fv3::dmin(-qa, fv3::dmax(qb, 0.0)); }

}  // namespace

extern "C" {

int fv3_sat_adjust(fv3_ctx *ctx, const fv3_sat_adjust_config *cf, double *te, double *qvapor, double *qliquid, double *qice,
                   double *qrain, double *qsnow, double *qgraupel, double *qcld, const double *hs_, const double *delp,
                   const double *delz, double *q_con, double *pt_, double *pkz, double *cappa, double zvir, double mdt,
                   int fast_mp_consv, int last_step, int kmp, void *stream) {
  const fv3_geom g = ctx->g;
  const fv3_grid m = ctx->m;
  const fv3_sat_adjust_config c = *cf;
  if (c.hydrostatic) {
    fv3::set_error("fv3_sat_adjust: hydrostatic is not implemented");
    return -1;
  }
  const int h = g.halo;
  const bool consv_te = fast_mp_consv != 0, last = last_step != 0;
  // conversion factors of SatAdjust3d.__call__ (:1023-1052), host libm as math.exp in the reference
  const double sdt = 0.5 * mdt;
  const double fac_i2s = 1.0 - exp(-mdt / c.tau_i2s), fac_v2l = 1.0 - exp(-sdt / c.tau_v2l);
  const double fac_r2g = 1.0 - exp(-mdt / c.tau_r2g), fac_l2r = 1.0 - exp(-mdt / c.tau_l2r);
  double fac_l2v = 1.0 - exp(-sdt / c.tau_l2v);
  fac_l2v = c.sat_adj0 < fac_l2v ? c.sat_adj0 : fac_l2v;
  const double fac_imlt = 1.0 - exp(-sdt / c.tau_imlt), fac_smlt = 1.0 - exp(-mdt / c.tau_smlt);
  const double c_air = CV_AIR, c_vap = CV_VAP;
  const double d0_vap = c_vap - C_LIQ, lv00 = HLV - d0_vap * TICE;
  const bool do_qa = true;
  fv3::launch3d(ctx, (cudaStream_t)stream, h, h + g.nx, h, h + g.ny, kmp, g.nz, FV_LAMBDA(int s, int i, int j, int k) { FV_DEV_GM
    const int64_t o = O3(s, i, j, k), o2 = O2(s, i, j);
    double qv = qvapor[o], ql = qliquid[o], qi = qice[o], qr = qrain[o], qs = qsnow[o], qg = qgraupel[o];
    const double dp = delp[o], dz = delz[o];
    double q_liq = ql + qr;
    double q_sol = qi + qs + qg;
    double qpz = q_liq + q_sol;
    double pt1 = pt_[o] / ((1.0 + zvir * qv) * (1.0 - qpz));
    const double t0 = pt1;
    qpz = qpz + qv;
    const double den = -dp / (GRAV * dz);
    const double mc_air = (1.0 - qpz) * c_air;
    double cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
    double lhi = LI00 + DC_ICE * pt1, icp2 = lhi / cvm;
    double lhl, lcp2;
    auto upd_i = [&]() {
      lhi = LI00 + DC_ICE * pt1;
      icp2 = lhi / cvm;
    };
    auto upd = [&]() {
      lhl = lv00 + d0_vap * pt1;
      lcp2 = lhl / cvm;
      upd_i();
    };
    double te0 = 0.0;
    if (consv_te) te0 = -cvm * t0;
    // fix negative cloud ice with snow
    if (qi < 0.0) {
      qs = qs + qi;
      qi = 0.0;
    }
    // melt_cloud_ice (:185-196)
    if (qi > 1.0e-8 && pt1 > TICE) {
      const double factmp = fac_imlt * (pt1 - TICE) / icp2;
      double sink = qi < factmp ? qi : factmp;
      qi = qi - sink;
      ql = ql + sink;
      q_liq = q_liq + sink;
      q_sol = q_sol - sink;
      cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
      sink = -sink;
      pt1 = pt1 + sink * lhi / cvm;
    }
    upd_i();
    // fix_negative_snow (:204-212)
    if (qs < 0.0) {
      qg = qg + qs;
      qs = 0.0;
    } else if (qg < 0.0) {
      const double tmp = minmax_tmp_h20(qg, qs);
      qg = qg + tmp;
      qs = qs - tmp;
    }
    // fix_negative_cloud_water (:216-226)
    if (ql < 0.0) {
      const double tmp = minmax_tmp_h20(ql, qr);
      ql = ql + tmp;
      qr = qr - tmp;
    } else if (qr < 0.0) {
      const double tmp = minmax_tmp_h20(qr, ql);
      ql = ql - tmp;
      qr = qr + tmp;
    }
    // complete_freezing below -48 C (:230-241)
    {
      const double dtmp = TICE - 48.0 - pt1;
      if (ql > 0.0 && dtmp > 0.0) {
        const double sink = fv3::dmin(ql, dtmp / icp2);
        ql = ql - sink;
        qi = qi + sink;
        q_liq = q_liq - sink;
        q_sol = q_sol + sink;
        cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
        pt1 = pt1 + sink * lhi / cvm;
      }
    }
    double wqsat, dq2dt;
    wqs2<true>(pt1, den, wqsat, dq2dt);
    upd();
    double tcp3 = lcp2 + icp2 * fv3::dmin(1.0, dimf(TICE, pt1) / 48.0);
    double dq0 = (qv - wqsat) / (1.0 + tcp3 * dq2dt);
    double src;
    // ql_evaporation (:421-425)
    auto evap = [&]() {
      const double factor = -fv3::dmin(1.0, fac_l2v * 10.0 * (1.0 - qv / wqsat));
      return -fv3::dmin(ql, factor * dq0);
    };
    if (dq0 > 0)
      src = fv3::dmin(c.sat_adj0 * dq0, fv3::dmax(c.ql_gen - ql, fac_v2l * dq0));
    else
      src = evap();
    // wqsat_correct (:428-435)
    auto correct = [&]() {
      qv = qv - src;
      ql = ql + src;
      q_liq = q_liq + src;
      cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
      pt1 = pt1 + src * lhl / cvm;
    };
    correct();
    upd();
    tcp3 = lcp2 + icp2 * fv3::dmin(1.0, dimf(TICE, pt1) / 48.0);
    if (last) {
      wqs2<true>(pt1, den, wqsat, dq2dt);
      dq0 = (qv - wqsat) / (1.0 + tcp3 * dq2dt);
      if (dq0 > 0)
        src = dq0;
      else
        src = evap();
      correct();
      upd();
    }
    // homogenous_freezing (:245-258)
    {
      const double dtmp = T_WFR - pt1;
      if (ql > 0.0 && dtmp > 0.0) {
        double sink = fv3::dmin(ql, dtmp / icp2);
        sink = fv3::dmin(sink, ql * dtmp * 0.125);
        ql = ql - sink;
        qi = qi + sink;
        q_liq = q_liq - sink;
        q_sol = q_sol + sink;
        cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
        pt1 = pt1 + sink * lhi / cvm;
      }
    }
    upd_i();
    const double exptc = exp(0.66 * (TICE0 - pt1));
    // heterogeneous_freezing, Bigg mechanism (:262-277)
    {
      const double tc = TICE0 - pt1;
      if (ql > 0.0 && tc > 0.0) {
        double sink = 3.3333e-10 * mdt * (exptc - 1.0) * den * (ql * ql);
        sink = fv3::dmin(ql, sink);
        sink = fv3::dmin(sink, tc / icp2);
        ql = ql - sink;
        qi = qi + sink;
        q_liq = q_liq - sink;
        q_sol = q_sol + sink;
        cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
        pt1 = pt1 + sink * lhi / cvm;
      }
    }
    upd_i();
    // make_graupel (:281-294)
    {
      const double dtmp = (TICE - 0.1) - pt1;
      if (qr > 1e-7 && dtmp > 0.0) {
        const double rainfac = (dtmp * 0.025) * (dtmp * 0.025);
        const double tmp = 1.0 < rainfac ? qr : rainfac * qr;
        const double sink = fv3::dmin(tmp, fac_r2g * dtmp / icp2);
        qr = qr - sink;
        qg = qg + sink;
        q_liq = q_liq - sink;
        q_sol = q_sol + sink;
        cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
        pt1 = pt1 + sink * lhi / cvm;
      }
    }
    upd_i();
    // melt_snow (:298-318)
    {
      const double dtmp = pt1 - (TICE + 0.1);
      const double dimqs = dimf(c.qs_mlt, ql);
      if (qs > 1e-7 && dtmp > 0.0) {
        const double snowfac = (dtmp * 0.1) * (dtmp * 0.1);
        double tmp = 1.0 < snowfac ? qs : snowfac * qs;
        const double sink = fv3::dmin(tmp, fac_smlt * dtmp / icp2);
        tmp = fv3::dmin(sink, dimqs);
        qs = qs - sink;
        ql = ql + tmp;
        qr = qr + sink - tmp;
        q_liq = q_liq + sink;
        q_sol = q_sol - sink;
        cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
        pt1 = pt1 - sink * lhi / cvm;
      }
    }
    // autoconversion_cloud_to_rain (:322-327)
    if (ql > c.ql0_max) {
      const double sink = fac_l2r * (ql - c.ql0_max);
      qr = qr + sink;
      ql = ql - sink;
    }
    double iqs2, dqsdt;
    wqs2<false>(pt1, den, iqs2, dqsdt);
    const double expsubl = exp(0.875 * log(qi * den));
    upd();
    const double tcp2 = lcp2 + icp2;
    const double adj_fac = last ? 1.0 : c.sat_adj0;
    // sublimation (:331-392)
    {
      double s2 = 0.0;
      if (pt1 < c.t_sub) {
        s2 = dimf(qv, 1e-6);
      } else if (pt1 < TICE0) {
        const double dq = qv - iqs2;
        const double sink = adj_fac * dq / (1.0 + tcp2 * dqsdt);
        double pidep = 0.0;
        if (qi > 1.0e-8)
          pidep = sdt * dq * 349138.78 * expsubl / (iqs2 * den * LAT2 / (0.0243 * RVGAS * (pt1 * pt1)) + 4.42478e4);
        if (dq > 0.0) {
          const double tmp = TICE - pt1;
          const double qi_crt = c.qi_lim < 0.1 * tmp ? c.qi_gen * c.qi_lim / den : c.qi_gen * 0.1 * tmp / den;
          const double maxtmp = qi_crt - qi > pidep ? qi_crt - qi : pidep;
          s2 = sink < maxtmp ? sink : maxtmp;
          s2 = s2 < tmp / tcp2 ? s2 : tmp / tcp2;
        } else {
          const double dimtmp = dimf(pt1, c.t_sub);
          pidep = 1.0 < (dimtmp * 0.2) ? pidep : pidep * dimtmp * 0.2;
          s2 = pidep > sink ? pidep : sink;
          s2 = s2 > -qi ? s2 : -qi;
        }
      }
      qv = qv - s2;
      qi = qi + s2;
      q_sol = q_sol + s2;
      cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol);
      const double lh = lhl + lhi;
      pt1 = pt1 + s2 * lh / cvm;
    }
    // virtual temperature updated
    const double qc = q_liq + q_sol;
    double tmp = 1.0 + zvir * qv;
    const double ptn = pt1 * tmp * (1.0 - qc);
    tmp *= RDGAS;
    const double cap = tmp / (tmp + cvm);
    // fix negative graupel with available cloud ice
    if (qg < 0) {
      const double mintmp = fv3::dmin(-qg, fv3::dmax(0.0, qi));
      qg = qg + mintmp;
      qi = qi - mintmp;
    }
    // autoconversion from cloud ice to snow
    const double qim = c.qi0_max / den;
    if (qi > qim) {
      const double sink = fac_i2s * (qi - qim);
      qi = qi - sink;
      qs = qs + sink;
    }
    if (consv_te) te[o] = dp * (te0 + cvm * pt1);
    cvm = mc_air + (qv + q_liq + q_sol) * c_vap;
    upd();
    // cloud fraction (:789-877)
    if (do_qa && last) {
      if (c.rad_snow)
        q_sol = c.rad_graupel ? qi + qs + qg : qi + qs;
      else
        q_sol = qi;
      q_liq = c.rad_rain ? ql + qr : ql;
      const double q_cond = q_sol + q_liq;
      const double tin = c.tintqs ? pt1 : pt1 - (lcp2 * q_cond + icp2 * q_sol);
      const double ap1 = ap1_for_wqs2(tin), it = floor(ap1);
      const double w1 = wqs1<true>(it, ap1, tin, den), i1 = wqs1<false>(it, ap1, tin, den);
      double qstar;
      if (tin < T_WFR) {
        qstar = i1;
      } else if (tin >= TICE) {
        qstar = w1;
      } else {
        const double rqi = q_cond > 1e-6 ? q_sol / q_cond : (TICE - tin) / (TICE - T_WFR);
        qstar = rqi * i1 + (1.0 - rqi) * w1;
      }
      const double mindw = fv3::dmin(1.0, fabs(hs_[o2]) / (10.0 * GRAV));
      const double dw = c.dw_ocean + (c.dw_land - c.dw_ocean) * mindw;
      const double dbl_sqrt_area = dw * sqrt(sqrt(m.area_64[o2]) / 100.0e3);
      const double hvar = fv3::dmin(0.2, fv3::dmax(0.01, dbl_sqrt_area));
      const double rh = qpz / qstar;
      double qa = 0.0;
      if (rh > 0.75 && qpz > 1.0e-8) {
        const double dq = hvar * qpz;
        const double q_plus = qpz + dq, q_minus = qpz - dq;
        if (c.icloud_f == 2) {
          if (qpz > qstar)
            qa = 1.0;
          else if (qstar < q_plus && q_cond > 1.0e-8) {
            const double r = (q_plus - qstar) / dq;
            qa = fv3::dmin(1.0, r * r);
          } else
            qa = 0.0;
        } else {
          if (qstar < q_minus) {
            qa = 1.0;
          } else {
            if (qstar < q_plus)
              qa = c.icloud_f == 0 ? (q_plus - qstar) / (dq + dq) : (q_plus - qstar) / (2.0 * dq * (1.0 - q_cond));
            else
              qa = 0.0;
            if (q_cond > 1.0e-8) qa = fv3::dmax(c.cld_min, qa);
            qa = fv3::dmin(1.0, qa);
          }
        }
      }
      qcld[o] = qa;
    }
    qvapor[o] = qv;
    qliquid[o] = ql;
    qice[o] = qi;
    qrain[o] = qr;
    qsnow[o] = qs;
    qgraupel[o] = qg;
    q_con[o] = qc;
    pt_[o] = ptn;
    cappa[o] = cap;
    // compute_pkz_func (moist_cv.py:125-127)
    pkz[o] = exp(cap * log(RDG * dp / dz * ptn));
  });
  return fv3::check_launch("fv3_sat_adjust");
}

}  // extern "C"
