"""Oracle: AdjustNegativeTracerMixingRatio (test infrastructure) — Python restatement, one column at a time, of
fv3core/pace/fv3core/stencils/neg_adj3.py: fix_neg_water (:98-143) with fix_negative_ice (:15-54) and
fix_negative_liq (:57-95), fillq (:146-175), fix_water_vapor_down (:179-249), fix_neg_cloud (:252-278), called in the
order of __call__ (:377-420).  Non-hydrostatic constants (:352-353)."""
import numpy as np

from .constants import C_ICE, C_LIQ, CV_AIR, CV_VAP, TICE

HLV, HLF = 2.5e6, 3.3358e5
DC_ICE = C_LIQ - C_ICE
LI0 = HLF - DC_ICE * TICE
D0_VAP = CV_VAP - C_LIQ
LV00 = HLV - D0_VAP * TICE


def _fix_negative_ice(qv, qi, qs, qg, qr, ql, pt, lcpk, icpk):
    qsum = qi + qs
    if qsum > 0.0:
        if qi < 0.0:
            qi, qs = 0.0, qsum
        elif qs < 0.0:
            qs, qi = 0.0, qsum
    else:
        qi, qs = 0.0, 0.0
        qg = qg + qsum
    if qg < 0.0:
        dq = qs if qs < -qg else -qg
        qs, qg = qs - dq, qg + dq
        if qg < 0.0:
            dq = qi if qi < -qg else -qg
            qi, qg = qi - dq, qg + dq
    if qg < 0.0 and qr > 0.0:
        dq = qr if qr < -qg else -qg
        qg, ql, pt = qg + dq, ql - dq, pt + dq * icpk
    if qg < 0.0 and ql > 0.0:
        dq = ql if ql < -qg else -qg
        qg, ql, pt = qg + dq, ql - dq, pt + dq * icpk
    if qg < 0.0 and qv > 0.0:
        dq = 0.999 * qv if 0.999 * qv < -qg else -qg
        qg, qv, pt = qg + dq, qv - dq, pt + dq * (icpk + lcpk)
    return qv, qi, qs, qg, qr, ql, pt


def _fix_negative_liq(qv, qi, qs, qg, qr, ql, pt, lcpk, icpk):
    qsum = ql + qr
    pos_qg = 0.0 if 0.0 > qg else qg
    if qsum > 0.0:
        if qr < 0.0:
            qr, ql = 0.0, qsum
        elif ql < 0.0:
            ql, qr = 0.0, qsum
    else:
        ql = 0.0
        qr_tmp = qsum
        dq = pos_qg if pos_qg < -qr_tmp else -qr_tmp
        qr_tmp, qg, pt = qr_tmp + dq, qg - dq, pt - dq * icpk
        if qr < 0.0:
            dq = qi + qs if (qi + qs) < -qr_tmp else -qr_tmp
            qr_tmp = qr_tmp + dq
            dq1 = dq if dq < qs else qs
            qs = qs - dq1
            qi = qi + dq1 - dq
            pt = pt - dq * icpk
        qr = qr_tmp
        if qr < 0.0 and qv > 0.0:
            dq = 0.999 * qv if 0.999 * qv < -qr else -qr
            qv, qr, pt = qv - dq, qr + dq, pt + dq * lcpk
    return qv, qi, qs, qg, qr, ql, pt


def _fillq(q, dp):
    km = len(q)
    sum1 = 0.0
    sum2 = 0.0
    for k in range(km):
        if q[k] > 0:
            sum1 = sum1 + q[k] * dp[k]
    for k in range(km - 1, -1, -1):
        if q[k] < 0.0 and sum1 >= 0:
            dq = sum1 if sum1 < -q[k] * dp[k] else -q[k] * dp[k]
            sum1, sum2 = sum1 - dq, sum2 + dq
            q[k] = q[k] + dq / dp[k]
    for k in range(km - 1, -1, -1):
        if q[k] > 0.0 and sum1 >= 1e-12 and sum2 > 0:
            dq = sum2 if sum2 < q[k] * dp[k] else q[k] * dp[k]
            sum2 = sum2 - dq
            q[k] = q[k] - dq / dp[k]


def _fix_water_vapor_down(q, dp):
    km = len(q)
    upper_fix = np.zeros(km)
    lower_fix = np.zeros(km)
    if q[0] < 0:
        q[1] = q[1] + q[0] * dp[0] / dp[1]
    if q[0] < 0.0:
        q[0] = 0.0
    for k in range(1, km - 1):
        dq = q[k - 1] * dp[k - 1]
        if lower_fix[k - 1] != 0:
            q[k] += lower_fix[k - 1] / dp[k]
        if q[k] < 0 and q[k - 1] > 0:
            dq = dq if dq < -q[k] * dp[k] else -q[k] * dp[k]
            upper_fix[k] = dq
            q[k] += dq / dp[k]
        if q[k] < 0:
            lower_fix[k] = q[k] * dp[k]
            q[k] = 0
    for k in range(km - 2):
        if upper_fix[k + 1] != 0:
            q[k] = q[k] - upper_fix[k + 1] / dp[k]
    kb = km - 1
    if lower_fix[kb - 1] > 0:
        q[kb] = q[kb] + lower_fix[kb] / dp[kb]
    upper_fix[kb] = q[kb]
    dp_bottom = dp[kb]
    for k in range(km - 2, -1, -1):
        dq = q[k] * dp[k]
        if upper_fix[k + 1] < 0 and q[k] > 0:
            if dq >= -upper_fix[k + 1] * dp_bottom:
                dq = -upper_fix[k + 1] * dp_bottom
            q[k] = q[k] - dq / dp[k]
            upper_fix[k] = upper_fix[k + 1] + dq / dp_bottom
        else:
            upper_fix[k] = upper_fix[k + 1]
    q[kb] = upper_fix[0]


def _fix_neg_cloud(q, dp):
    km = len(q)
    for k in range(1, km - 1):
        if q[k - 1] < 0.0:
            q[k] = q[k] + q[k - 1] * dp[k - 1] / dp[k]
    for k in range(1, km - 1):
        if q[k] < 0.0:
            q[k] = 0.0
    k = km - 2
    if q[k + 1] < 0.0 and q[k] > 0:
        dq = -q[k] * dp[k] if -q[k] * dp[k] < q[k + 1] * dp[k + 1] else q[k + 1] * dp[k + 1]
        q[k] = q[k] - dq / dp[k]
    k = km - 1
    if q[k] < 0 and q[k - 1] > 0.0:
        dq = -q[k] * dp[k] if -q[k] * dp[k] < q[k - 1] * dp[k - 1] else q[k - 1] * dp[k - 1]
        q[k] = q[k] + dq / dp[k]
        q[k] = 0.0 if 0.0 > q[k] else q[k]


def neg_adj3(qvapor, qliquid, qrain, qsnow, qice, qgraupel, qcld, pt, delp):
    """All arrays [ni, nj, km] (compute domain); the first eight are adjusted in place."""
    ni, nj, km = pt.shape
    for i in range(ni):
        for j in range(nj):
            for k in range(km):
                qv, ql, qr = qvapor[i, j, k], qliquid[i, j, k], qrain[i, j, k]
                qs, qi, qg, t = qsnow[i, j, k], qice[i, j, k], qgraupel[i, j, k], pt[i, j, k]
                q_liq = 0.0 if 0.0 > ql + qr else ql + qr
                q_sol = 0.0 if 0.0 > qi + qs else qi + qs
                cpm = (1.0 - (qv + q_liq + q_sol)) * CV_AIR + qv * CV_VAP + q_liq * C_LIQ + q_sol * C_ICE
                lcpk = (LV00 + D0_VAP * t) / cpm
                icpk = (LI0 + DC_ICE * t) / cpm
                qv, qi, qs, qg, qr, ql, t = _fix_negative_ice(qv, qi, qs, qg, qr, ql, t, lcpk, icpk)
                qv, qi, qs, qg, qr, ql, t = _fix_negative_liq(qv, qi, qs, qg, qr, ql, t, lcpk, icpk)
                qvapor[i, j, k], qliquid[i, j, k], qrain[i, j, k] = qv, ql, qr
                qsnow[i, j, k], qice[i, j, k], qgraupel[i, j, k], pt[i, j, k] = qs, qi, qg, t
            dp = delp[i, j, :]
            _fillq(qgraupel[i, j, :], dp)
            _fillq(qrain[i, j, :], dp)
            _fix_water_vapor_down(qvapor[i, j, :], dp)
            _fix_neg_cloud(qcld[i, j, :], dp)
