"""numpy restatement of the fast saturation adjustment — TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows fv3core/pace/fv3core/stencils/saturation_adjustment.py statement by statement under GT4Py's PARALLEL
semantics (every `if` is a mask, both sides evaluated): the `satadjust` stencil (:561-943), its gtscript functions
(:30-560) and the conversion factors of `SatAdjust3d.__call__` (:1023-1052), non-hydrostatic branch.  Checked against
the reference's own SatAdjust3d inputs / outputs (tests/golden/c12satk2_step/stage_rank0) by tests/test_sat_adjust.py.
"""
import math

import numpy as np

GRAV, RDGAS, RVGAS, HLV, HLF, CP_AIR = 9.80665, 287.05, 461.50, 2.5e6, 3.3358e5, 1004.6
CV_AIR, RDG, CV_VAP, C_ICE, C_LIQ = CP_AIR - RDGAS, -RDGAS / GRAV, 3.0 * RVGAS, 1972.0, 4.1855e3
CP_VAP, TICE = 4.0 * RVGAS, 273.16
DC_ICE, DC_VAP = C_LIQ - C_ICE, CP_VAP - C_LIQ
D2ICE = DC_VAP + DC_ICE
LV0, LI00 = HLV - DC_VAP * TICE, HLF - DC_ICE * TICE
LI2, E00, T_WFR, TICE0, T_SAT_MIN = LV0 + LI00, 611.21, TICE - 40.0, TICE - 0.01, TICE - 160.0
LAT2 = (HLV + HLF) ** 2
DELT, QS_LENGTH = 0.1, 2621

DEFAULTS = dict(rad_snow=True, rad_rain=True, rad_graupel=True, tintqs=False, sat_adj0=0.90, ql_gen=1.0e-3, qs_mlt=1.0e-6,
                ql0_max=2.0e-3, t_sub=184.0, qi_gen=1.82e-6, qi_lim=1.0, qi0_max=1.0e-4, dw_ocean=0.10, dw_land=0.15,
                icloud_f=0, cld_min=0.05, tau_i2s=1000.0, tau_v2l=90.0, tau_r2g=900.0, tau_l2r=900.0, tau_l2v=300.0,
                tau_imlt=600.0, tau_smlt=900.0)


def dim(a, b):
    return np.where(a - b > 0, a - b, 0.0)


def _oneline(dhc, lhc, tem):
    return E00 * np.exp((dhc * np.log(tem / TICE) + (tem - TICE) / (tem * TICE) * lhc) / RVGAS)


def table_vapor(tem):
    return _oneline(DC_VAP, LV0, tem)


def table_ice(tem):
    return _oneline(D2ICE, LI2, tem)


def tem_lower(i):
    return T_SAT_MIN + DELT * i


def tem_upper(i):
    return 253.16 + DELT * i


def qs_table2(i):  # :93-127
    i = np.asarray(i, dtype=np.float64)
    tem0 = tem_lower(i)
    t2 = np.where(i < 1600, table_ice(tem0), table_vapor(tem0))
    tu = tem_upper(i - 1400)
    table = (0.05 * (TICE - tu)) * table_ice(tem0) + (0.05 * (tu - 253.16)) * table_vapor(tu)
    t2 = np.where(i == 1599, 0.25 * (table_ice(tem_lower(1598.0)) + 2.0 * table + table_vapor(tem_lower(1600.0))), t2)
    t2 = np.where(i == 1600, 0.25 * (table_ice(tem_lower(1599.0)) + 2.0 * table_vapor(tu) + table_vapor(tem_lower(1601.0))), t2)
    return t2


def qs_tablew(i):
    return table_vapor(tem_lower(np.asarray(i, dtype=np.float64)))


def des2_table(i):  # :146-154 with des_end :133-139
    t = qs_table2(i)
    d = np.maximum(0.0, qs_table2(i + 1) - t)
    return np.where(i == QS_LENGTH - 1, np.maximum(0.0, t - qs_table2(i - 1)), d)


def desw_table(i):
    t = qs_tablew(i)
    d = np.maximum(0.0, qs_tablew(i + 1) - t)
    return np.where(i == QS_LENGTH - 1, np.maximum(0.0, t - qs_table2(i - 1)), d)


def ap1_for_wqs2(ta):
    return np.minimum(10.0 * dim(ta, T_SAT_MIN) + 1.0, QS_LENGTH) - 1


def wqs2(ta, den, water):  # :468-493
    ap1 = ap1_for_wqs2(ta)
    it, it2 = np.floor(ap1), np.floor(ap1 - 0.5)
    tab, des = (qs_tablew, desw_table) if water else (qs_table2, des2_table)
    es = tab(it) + (ap1 - it) * des(it)
    denom = RVGAS * ta * den
    dqdt = 10.0 * (des(it2) + (ap1 - it2) * (des(it2 + 1) - des(it2)))
    return es / denom, dqdt / denom


def wqs1(it, ap1, ta, den, water):
    tab, des = (qs_tablew, desw_table) if water else (qs_table2, des2_table)
    return (tab(it) + (ap1 - it) * des(it)) / (RVGAS * ta * den)


def compute_cvm(mc_air, qv, c_vap, q_liq, q_sol):
    return mc_air + qv * c_vap + q_liq * C_LIQ + q_sol * C_ICE


def sat_adjust(a, area, hs, r_vir, mdt, fast_mp_consv, last_step, kmp, nz, config=None):
    """In-place on the dict `a` of [i, j, k] arrays (compute-domain slices): qvapor, qliquid, qice, qrain, qsnow, qgraupel,
    qcld, delp, delz, q_con, pt, pkz, cappa, te; area / hs are [i, j]."""
    c = dict(DEFAULTS)
    c.update(config or {})
    K = slice(kmp, nz)
    g = lambda n: a[n][:, :, K].copy()  # noqa: E731
    qv, ql, qi, qr, qs, qg = g("qvapor"), g("qliquid"), g("qice"), g("qrain"), g("qsnow"), g("qgraupel")
    dp, dz, pt = g("delp"), g("delz"), g("pt")
    sdt = 0.5 * mdt
    fac_i2s, fac_v2l = 1.0 - math.exp(-mdt / c["tau_i2s"]), 1.0 - math.exp(-sdt / c["tau_v2l"])
    fac_r2g, fac_l2r = 1.0 - math.exp(-mdt / c["tau_r2g"]), 1.0 - math.exp(-mdt / c["tau_l2r"])
    fac_l2v = min(c["sat_adj0"], 1.0 - math.exp(-sdt / c["tau_l2v"]))
    fac_imlt, fac_smlt = 1.0 - math.exp(-sdt / c["tau_imlt"]), 1.0 - math.exp(-mdt / c["tau_smlt"])
    c_air, c_vap = CV_AIR, CV_VAP
    d0_vap = c_vap - C_LIQ
    lv00 = HLV - d0_vap * TICE
    W = np.where
    with np.errstate(all="ignore"):
        q_liq = ql + qr
        q_sol = qi + qs + qg
        qpz = q_liq + q_sol
        pt1 = pt / ((1.0 + r_vir * qv) * (1.0 - qpz))
        t0 = pt1
        qpz = qpz + qv
        den = -dp / (GRAV * dz)
        mc_air = (1.0 - qpz) * c_air
        cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol)
        lhi = LI00 + DC_ICE * pt1
        icp2 = lhi / cvm
        te0 = -cvm * t0
        m = qi < 0.0
        qs = W(m, qs + qi, qs)
        qi = W(m, 0.0, qi)
        # melt_cloud_ice
        m = (qi > 1.0e-8) & (pt1 > TICE)
        factmp = fac_imlt * (pt1 - TICE) / icp2
        sink = W(qi < factmp, qi, factmp)
        qi, ql = W(m, qi - sink, qi), W(m, ql + sink, ql)
        q_liq, q_sol = W(m, q_liq + sink, q_liq), W(m, q_sol - sink, q_sol)
        cvm = W(m, compute_cvm(mc_air, qv, c_vap, q_liq, q_sol), cvm)
        pt1 = W(m, pt1 + (-sink) * lhi / cvm, pt1)
        lhi = LI00 + DC_ICE * pt1
        icp2 = lhi / cvm
        # fix_negative_snow
        m1 = qs < 0.0
        m2 = ~m1 & (qg < 0.0)
        tmp = np.minimum(-qg, np.maximum(qs, 0.0))
        qg_n = W(m1, qg + qs, W(m2, qg + tmp, qg))
        qs = W(m1, 0.0, W(m2, qs - tmp, qs))
        qg = qg_n
        # fix_negative_cloud_water
        m1 = ql < 0.0
        m2 = ~m1 & (qr < 0.0)
        t1 = np.minimum(-ql, np.maximum(qr, 0.0))
        t2 = np.minimum(-qr, np.maximum(ql, 0.0))
        ql_n = W(m1, ql + t1, W(m2, ql - t2, ql))
        qr = W(m1, qr - t1, W(m2, qr + t2, qr))
        ql = ql_n

        def freeze(m, sink, ql, qi, q_liq, q_sol, cvm, pt1):
            ql, qi = W(m, ql - sink, ql), W(m, qi + sink, qi)
            q_liq, q_sol = W(m, q_liq - sink, q_liq), W(m, q_sol + sink, q_sol)
            cvm = W(m, compute_cvm(mc_air, qv, c_vap, q_liq, q_sol), cvm)
            pt1 = W(m, pt1 + sink * lhi / cvm, pt1)
            return ql, qi, q_liq, q_sol, cvm, pt1

        # complete_freezing
        dtmp = TICE - 48.0 - pt1
        ql, qi, q_liq, q_sol, cvm, pt1 = freeze((ql > 0.0) & (dtmp > 0.0), np.minimum(ql, dtmp / icp2), ql, qi, q_liq, q_sol, cvm, pt1)
        wqsat, dq2dt = wqs2(pt1, den, True)

        def upd(pt1, cvm):
            lhl = lv00 + d0_vap * pt1
            lhi = LI00 + DC_ICE * pt1
            return lhl, lhi, lhl / cvm, lhi / cvm

        lhl, lhi, lcp2, icp2 = upd(pt1, cvm)
        tcp3 = lcp2 + icp2 * np.minimum(1.0, dim(TICE, pt1) / 48.0)
        dq0 = (qv - wqsat) / (1.0 + tcp3 * dq2dt)

        def evap(wqsat, qv, ql, dq0):
            factor = -np.minimum(1, fac_l2v * 10.0 * (1.0 - qv / wqsat))
            return -np.minimum(ql, factor * dq0)

        src = W(dq0 > 0, np.minimum(c["sat_adj0"] * dq0, np.maximum(c["ql_gen"] - ql, fac_v2l * dq0)), evap(wqsat, qv, ql, dq0))

        def correct(src, pt1, lhl, qv, ql, q_liq):
            qv, ql, q_liq = qv - src, ql + src, q_liq + src
            cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol)
            return qv, ql, q_liq, cvm, pt1 + src * lhl / cvm

        qv, ql, q_liq, cvm, pt1 = correct(src, pt1, lhl, qv, ql, q_liq)
        lhl, lhi, lcp2, icp2 = upd(pt1, cvm)
        tcp3 = lcp2 + icp2 * np.minimum(1.0, dim(TICE, pt1) / 48.0)
        if last_step:
            wqsat, dq2dt = wqs2(pt1, den, True)
            dq0 = (qv - wqsat) / (1.0 + tcp3 * dq2dt)
            src = W(dq0 > 0, dq0, evap(wqsat, qv, ql, dq0))
            qv, ql, q_liq, cvm, pt1 = correct(src, pt1, lhl, qv, ql, q_liq)
            lhl, lhi, lcp2, icp2 = upd(pt1, cvm)
        # homogenous_freezing
        dtmp = T_WFR - pt1
        sink = np.minimum(np.minimum(ql, dtmp / icp2), ql * dtmp * 0.125)
        ql, qi, q_liq, q_sol, cvm, pt1 = freeze((ql > 0.0) & (dtmp > 0.0), sink, ql, qi, q_liq, q_sol, cvm, pt1)
        lhi = LI00 + DC_ICE * pt1
        icp2 = lhi / cvm
        exptc = np.exp(0.66 * (TICE0 - pt1))
        # heterogeneous_freezing
        tc = TICE0 - pt1
        sink = 3.3333e-10 * mdt * (exptc - 1.0) * den * ql ** 2
        sink = np.minimum(np.minimum(ql, sink), tc / icp2)
        ql, qi, q_liq, q_sol, cvm, pt1 = freeze((ql > 0.0) & (tc > 0.0), sink, ql, qi, q_liq, q_sol, cvm, pt1)
        lhi = LI00 + DC_ICE * pt1
        icp2 = lhi / cvm
        # make_graupel
        dtmp = (TICE - 0.1) - pt1
        m = (qr > 1e-7) & (dtmp > 0.0)
        rainfac = (dtmp * 0.025) ** 2
        sink = np.minimum(W(1.0 < rainfac, qr, rainfac * qr), fac_r2g * dtmp / icp2)
        qr, qg = W(m, qr - sink, qr), W(m, qg + sink, qg)
        q_liq, q_sol = W(m, q_liq - sink, q_liq), W(m, q_sol + sink, q_sol)
        cvm = W(m, compute_cvm(mc_air, qv, c_vap, q_liq, q_sol), cvm)
        pt1 = W(m, pt1 + sink * lhi / cvm, pt1)
        lhi = LI00 + DC_ICE * pt1
        icp2 = lhi / cvm
        # melt_snow
        dtmp = pt1 - (TICE + 0.1)
        dimqs = dim(c["qs_mlt"], ql)
        m = (qs > 1e-7) & (dtmp > 0.0)
        snowfac = (dtmp * 0.1) ** 2
        sink = np.minimum(W(1.0 < snowfac, qs, snowfac * qs), fac_smlt * dtmp / icp2)
        tmp = np.minimum(sink, dimqs)
        qs, ql, qr = W(m, qs - sink, qs), W(m, ql + tmp, ql), W(m, qr + sink - tmp, qr)
        q_liq, q_sol = W(m, q_liq + sink, q_liq), W(m, q_sol - sink, q_sol)
        cvm = W(m, compute_cvm(mc_air, qv, c_vap, q_liq, q_sol), cvm)
        pt1 = W(m, pt1 - sink * lhi / cvm, pt1)
        # autoconversion_cloud_to_rain
        m = ql > c["ql0_max"]
        sink = fac_l2r * (ql - c["ql0_max"])
        qr, ql = W(m, qr + sink, qr), W(m, ql - sink, ql)
        iqs2, dqsdt = wqs2(pt1, den, False)
        expsubl = np.exp(0.875 * np.log(qi * den))
        lhl, lhi, lcp2, icp2 = upd(pt1, cvm)
        tcp2 = lcp2 + icp2
        adj_fac = 1.0 if last_step else c["sat_adj0"]
        # sublimation
        dq = qv - iqs2
        sink = adj_fac * dq / (1.0 + tcp2 * dqsdt)
        pidep = W(qi > 1.0e-8, sdt * dq * 349138.78 * expsubl / (iqs2 * den * LAT2 / (0.0243 * RVGAS * pt1 ** 2.0) + 4.42478e4), 0.0)
        tmp = TICE - pt1
        qi_crt = W(c["qi_lim"] < 0.1 * tmp, c["qi_gen"] * c["qi_lim"] / den, c["qi_gen"] * 0.1 * tmp / den)
        maxtmp = W(qi_crt - qi > pidep, qi_crt - qi, pidep)
        s_pos = W(sink < maxtmp, sink, maxtmp)
        s_pos = W(s_pos < tmp / tcp2, s_pos, tmp / tcp2)
        dimtmp = dim(pt1, c["t_sub"])
        pd2 = W(1.0 < (dimtmp * 0.2), pidep, pidep * dimtmp * 0.2)
        s_neg = W(pd2 > sink, pd2, sink)
        s_neg = W(s_neg > -qi, s_neg, -qi)
        src = W(pt1 < c["t_sub"], dim(qv, 1e-6), W(pt1 < TICE0, W(dq > 0.0, s_pos, s_neg), 0.0))
        qv, qi, q_sol = qv - src, qi + src, q_sol + src
        cvm = compute_cvm(mc_air, qv, c_vap, q_liq, q_sol)
        pt1 = pt1 + src * (lhl + lhi) / cvm
        q_con = q_liq + q_sol
        tmp = 1.0 + r_vir * qv
        pt_new = pt1 * tmp * (1.0 - q_con)
        tmp = tmp * RDGAS
        cappa = tmp / (tmp + cvm)
        m = qg < 0
        mintmp = np.minimum(-qg, np.maximum(0.0, qi))
        qg, qi = W(m, qg + mintmp, qg), W(m, qi - mintmp, qi)
        qim = c["qi0_max"] / den
        m = qi > qim
        sink = fac_i2s * (qi - qim)
        qi, qs = W(m, qi - sink, qi), W(m, qs + sink, qs)
        if fast_mp_consv:
            a["te"][:, :, K] = dp * (te0 + cvm * pt1)
        cvm = mc_air + (qv + q_liq + q_sol) * c_vap
        lhl, lhi, lcp2, icp2 = upd(pt1, cvm)
        if last_step:
            q_sol = (qi + qs + qg if c["rad_graupel"] else qi + qs) if c["rad_snow"] else qi
            q_liq = ql + qr if c["rad_rain"] else ql
            q_cond = q_sol + q_liq
            tin = pt1 if c["tintqs"] else pt1 - (lcp2 * q_cond + icp2 * q_sol)
            ap1 = ap1_for_wqs2(tin)
            it = np.floor(ap1)
            w1, i1 = wqs1(it, ap1, tin, den, True), wqs1(it, ap1, tin, den, False)
            rqi = W(q_cond > 1e-6, q_sol / q_cond, (TICE - tin) / (TICE - T_WFR))
            qstar = W(tin < T_WFR, i1, W(tin >= TICE, w1, rqi * i1 + (1.0 - rqi) * w1))
            mindw = np.minimum(1.0, np.abs(hs) / (10.0 * GRAV))[:, :, None]
            dw = c["dw_ocean"] + (c["dw_land"] - c["dw_ocean"]) * mindw
            hvar = np.minimum(0.2, np.maximum(0.01, dw * (area[:, :, None] ** 0.5 / 100.0e3) ** 0.5))
            rh = qpz / qstar
            dqv = hvar * qpz
            q_plus, q_minus = qpz + dqv, qpz - dqv
            if c["icloud_f"] == 2:
                qa = W(qpz > qstar, 1.0, W((qstar < q_plus) & (q_cond > 1.0e-8), np.minimum(1.0, ((q_plus - qstar) / dqv) ** 2), 0.0))
            else:
                part = (q_plus - qstar) / (dqv + dqv) if c["icloud_f"] == 0 else (q_plus - qstar) / (2.0 * dqv * (1.0 - q_cond))
                qa = W(qstar < q_plus, part, 0.0)
                qa = W(q_cond > 1.0e-8, np.maximum(c["cld_min"], qa), qa)
                qa = np.minimum(1, qa)
                qa = W(qstar < q_minus, 1.0, qa)
            a["qcld"][:, :, K] = W((rh > 0.75) & (qpz > 1.0e-8), qa, 0.0)
        for n, v in (("qvapor", qv), ("qliquid", ql), ("qice", qi), ("qrain", qr), ("qsnow", qs), ("qgraupel", qg),
                     ("q_con", q_con), ("pt", pt_new), ("cappa", cappa)):
            a[n][:, :, K] = v
        a["pkz"][:, :, K] = np.exp(cappa * np.log(RDG * dp / dz * pt_new))
