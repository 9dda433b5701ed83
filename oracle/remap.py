"""Oracle: vertical remapping pieces (test infrastructure) — column-by-column numpy / Python restatement of
MapSingle.__call__ (fv3core/pace/fv3core/stencils/map_single.py:147-200) = RemapProfile for kord 9
(remap_profile.py:150-563, 622-681) + lagrangian_contributions (map_single.py:21-81), and of
FillNegativeTracerValues (fillz.py:15-108).  Written from the reference's statements, one column at a time
(small cases only)."""
import numpy as np


def _posdef_iv1(a1, a2, a3, a4):
    da1 = a3 - a2
    da2 = da1 * da1
    a6da = a4 * da1
    if ((a1 - a2) * (a1 - a3)) >= 0.0:
        return a1, a1, 0.0
    if a6da < -1.0 * da2:
        a4 = 3.0 * (a2 - a1)
        return a2, a2 - a4, a4
    if a6da > da2:
        a4 = 3.0 * (a3 - a1)
        return a3 - a4, a3, a4
    return a2, a3, a4


def _constraint(a1, a2, a3, a4, extm):
    da1 = a3 - a2
    da2 = da1 * da1
    a6da = a4 * da1
    if extm:
        return a1, a1, 0.0
    if a6da < -da2:
        a4 = 3.0 * (a2 - a1)
        return a2, a2 - a4, a4
    if a6da > da2:
        a4 = 3.0 * (a3 - a1)
        return a3 - a4, a3, a4
    return a2, a3, a4


def _posdef_iv0(a1, a2, a3, a4):
    if a1 <= 0.0:
        return a1, a1, 0.0
    if abs(a3 - a2) < -a4 and (a1 + 0.25 * ((a3 - a2) * (a3 - a2)) / a4 + a4 * (1.0 / 12.0)) < 0.0:
        if a1 < a3 and a1 < a2:
            return a1, a1, 0.0
        if a3 > a2:
            a4 = 3.0 * (a2 - a1)
            return a2, a2 - a4, a4
        a4 = 3.0 * (a3 - a1)
        return a3 - a4, a3, a4
    return a2, a3, a4


def remap_profile_column(a1, delp, km, iv, qs, qmin):
    """RemapProfile.__call__ for kord 9 on one column: returns (a2, a3, a4) of every layer."""
    q = np.zeros(km + 1)
    gam = np.zeros(km + 1)
    if iv != -2:                                                   # set_initial_vals :150-213
        grat = delp[1] / delp[0]
        bet = grat * (grat + 0.5)
        q[0] = ((grat + grat) * (grat + 1.0) * a1[0] + a1[1]) / bet
        gam[0] = (1.0 + grat * (grat + 1.5)) / bet
        for k in range(1, km):
            d4 = delp[k - 1] / delp[k]
            bet = 2.0 + d4 + d4 - gam[k - 1]
            q[k] = (3.0 * (a1[k - 1] + d4 * a1[k]) - q[k - 1]) / bet
            gam[k] = d4 / bet
        d4 = delp[km - 2] / delp[km - 1]
        a_bot = 1.0 + d4 * (d4 + 1.5)
        q[km] = (2.0 * d4 * (d4 + 1.0) * a1[km - 1] + a1[km - 2] - a_bot * q[km - 1]) / (d4 * (d4 + 0.5) - a_bot * gam[km - 1])
        for k in range(km - 1, -1, -1):
            q[k] = q[k] - gam[k] * q[k + 1]
    else:                                                          # :214-250
        gr = np.zeros(km + 1)
        q[0] = 1.5 * a1[0]
        gam[1] = 0.5
        gr[1] = delp[0] / delp[1]
        q[1] = (3.0 * (a1[0] + a1[1]) - q[0]) / (2.0 + gr[1] + gr[1] - gam[1])
        for k in range(2, km):
            old_gr = delp[k - 2] / delp[k - 1]
            old_bet = 2.0 + old_gr + old_gr - gam[k - 1]
            gam[k] = old_gr / old_bet
            gr[k] = delp[k - 1] / delp[k]
        for k in range(2, km - 1):
            bet = 2.0 + gr[k] + gr[k] - gam[k]
            q[k] = (3.0 * (a1[k - 1] + a1[k]) - q[k - 1]) / bet
        q[km - 1] = (3.0 * (a1[km - 2] + a1[km - 1]) - gr[km - 1] * qs - q[km - 2]) / (2.0 + gr[km - 1] + gr[km - 1] - gam[km - 1])
        q[km] = qs
        for k in range(km - 2, -1, -1):
            q[k] = q[k] - gam[k + 1] * q[k + 1]
    for k in range(1, km):                                         # apply_constraints :253-337
        gam[k] = a1[k] - a1[k - 1]
    for k in range(1, km):
        tmp, tmp2 = max(a1[k - 1], a1[k]), min(a1[k - 1], a1[k])
        if k == 1 or k == km - 1:
            q[k] = min(q[k], tmp)
            q[k] = max(q[k], tmp2)
        elif gam[k - 1] * gam[k + 1] > 0:
            q[k] = min(q[k], tmp)
            q[k] = max(q[k], tmp2)
        elif gam[k - 1] > 0:
            q[k] = max(q[k], tmp2)
        else:
            q[k] = min(q[k], tmp)
            if iv == 0 and q[k] < 0.0:
                q[k] = 0.0
    a2 = q[:km].copy()
    a3 = q[1 : km + 1].copy()
    a4 = np.zeros(km)
    extm = np.zeros(km, dtype=bool)
    for k in range(1, km - 1):
        extm[k] = gam[k] * gam[k + 1] < 0.0
    if iv == 0 and a2[0] < 0.0:                                    # set_interpolation_coefficients :340-563
        a2[0] = 0.0
    if iv == -1 and a2[0] * a1[0] <= 0.0:
        a2[0] = 0.0
    for k in (0, 1):
        a4[k] = 3.0 * (2.0 * a1[k] - (a2[k] + a3[k]))
    a2[0], a3[0], a4[0] = _posdef_iv1(a1[0], a2[0], a3[0], a4[0])
    a2[1], a3[1], a4[1] = _constraint(a1[1], a2[1], a3[1], a4[1], extm[1])
    for k in range(2, km - 2):
        v1 = a1[k]
        pmp_1 = v1 - 2.0 * gam[k + 1]
        lac_1 = pmp_1 + 1.5 * gam[k + 2]
        pmp_2 = v1 + 2.0 * gam[k]
        lac_2 = pmp_2 - 1.5 * gam[k - 1]
        if (extm[k] and extm[k - 1]) or (extm[k] and extm[k + 1]) or (extm[k] and (qmin > 0.0 and v1 < qmin)):
            a2[k] = v1
            a3[k] = v1
            a4[k] = 0.0
        else:
            a4[k] = 6.0 * v1 - 3.0 * (a2[k] + a3[k])
            if abs(a4[k]) > abs(a2[k] - a3[k]):
                tmin, tmax = min(v1, pmp_1, lac_1), max(v1, pmp_1, lac_1)
                a2[k] = min(max(a2[k], tmin), tmax)
                tmin, tmax = min(v1, pmp_2, lac_2), max(v1, pmp_2, lac_2)
                a3[k] = min(max(a3[k], tmin), tmax)
                a4[k] = 6.0 * v1 - 3.0 * (a2[k] + a3[k])
        if iv == 0:
            a2[k], a3[k], a4[k] = _posdef_iv0(a1[k], a2[k], a3[k], a4[k])
    if iv == 0 and a3[km - 1] < 0.0:
        a3[km - 1] = 0.0
    if iv == -1 and a3[km - 1] * a1[km - 1] <= 0.0:
        a3[km - 1] = 0.0
    for k in (km - 2, km - 1):
        a4[k] = 3.0 * (2.0 * a1[k] - (a2[k] + a3[k]))
    a2[km - 2], a3[km - 2], a4[km - 2] = _constraint(a1[km - 2], a2[km - 2], a3[km - 2], a4[km - 2], extm[km - 2])
    a2[km - 1], a3[km - 1], a4[km - 1] = _posdef_iv1(a1[km - 1], a2[km - 1], a3[km - 1], a4[km - 1])
    return a2, a3, a4


def map_single(q1, pe1, pe2, qs, qmin, iv, nx, ny, km, i_extra=0, j_extra=0, halo=3):
    """MapSingle.__call__ (map_single.py:147-200), kord 9; q1 [i, j, k] remapped in place from pe1 to pe2 layers."""
    for i in range(halo, halo + nx + i_extra):
        for j in range(halo, halo + ny + j_extra):
            p1 = pe1[i, j, : km + 1]
            dp1 = p1[1:] - p1[:-1]
            a1 = q1[i, j, :km].copy()
            qsv = 0.0 if qs is None else float(qs[i, j] if qs.ndim == 2 else qs[i, j, 0])
            a2, a3, a4 = remap_profile_column(a1, dp1, km, iv, qsv, qmin)
            L = 0                                                  # lagrangian_contributions :21-81
            top = pe2[i, j, 0]
            for k in range(km):
                bot = pe2[i, j, k + 1]
                pl = (top - p1[L]) / dp1[L]
                if bot <= p1[L + 1]:
                    pr = (bot - p1[L]) / dp1[L]
                    out = a2[L] + 0.5 * (a4[L] + a3[L] - a2[L]) * (pr + pl) - a4[L] * 1.0 / 3.0 * (pr * (pr + pl) + pl * pl)
                else:
                    qsum = (p1[L + 1] - top) * (a2[L] + 0.5 * (a4[L] + a3[L] - a2[L]) * (1.0 + pl) - a4[L] * 1.0 / 3.0 * (1.0 + pl * (1.0 + pl)))
                    L += 1
                    while L + 1 <= km and p1[L + 1] < bot:
                        qsum += dp1[L] * a1[L]
                        L += 1
                    L = min(L, km - 1)
                    dp = bot - p1[L]
                    esl = dp / dp1[L]
                    qsum += dp * (a2[L] + 0.5 * esl * (a3[L] - a2[L] + a4[L] * (1.0 - (2.0 / 3.0) * esl)))
                    out = qsum / (bot - top)
                q1[i, j, k] = out
                top = bot


def fillz(q, dp, nx, ny, km, halo=3):
    """fix_tracer (fillz.py:15-108) on one tracer field, compute domain, in place."""
    for i in range(halo, halo + nx):
        for j in range(halo, halo + ny):
            qc = q[i, j, :km].copy()
            d = dp[i, j, :km]
            lower_fix = np.zeros(km)
            upper_fix = np.zeros(km)
            zfix = 0
            if qc[0] < 0.0:
                qc[1] = qc[1] + qc[0] * d[0] / d[1]
            if qc[0] < 0:
                qc[0] = 0
            dm = np.zeros(km)
            dm_pos = np.zeros(km)
            dm[0] = qc[0] * d[0]
            for k in range(1, km - 1):
                if lower_fix[k - 1] != 0.0:
                    qc[k] = qc[k] - (lower_fix[k - 1] / d[k])
                if qc[k] < 0.0:
                    zfix += 1
                    if qc[k - 1] > 0.0:
                        dq = min(qc[k - 1] * d[k - 1], -(qc[k] * d[k]))
                        qc[k] = qc[k] + dq / d[k]
                        upper_fix[k] = dq
                    if qc[k] < 0.0 and qc[k + 1] > 0.0:
                        dq = min(qc[k + 1] * d[k + 1], -(qc[k] * d[k]))
                        qc[k] = qc[k] + dq / d[k]
                        lower_fix[k] = dq
            for k in range(km - 1):
                if upper_fix[k + 1] != 0.0:
                    qc[k] = qc[k] - upper_fix[k + 1] / d[k]
                dm[k] = qc[k] * d[k]
                dm_pos[k] = max(dm[k], 0.0)
            k = km - 1
            if lower_fix[k - 1] != 0.0:
                qc[k] = qc[k] - (lower_fix[k - 1] / d[k])
            qup = qc[k - 1] * d[k - 1]
            qly = -qc[k] * d[k]
            dup = min(qup, qly)
            if qc[k] < 0.0 and qc[k - 1] > 0.0:
                zfix += 1
                qc[k] = qc[k] + (dup / d[k])
                upper_fix[k] = dup
            dm[k] = qc[k] * d[k]
            dm_pos[k] = max(dm[k], 0.0)
            k = km - 2
            if upper_fix[k + 1] != 0.0:
                qc[k] = qc[k] - (upper_fix[k + 1] / d[k])
                dm[k] = qc[k] * d[k]
                dm_pos[k] = max(dm[k], 0.0)
            sum0 = 0.0
            sum1 = 0.0
            for k in range(1, km):
                sum0 += dm[k]
                sum1 += dm_pos[k]
            fac = sum0 / sum1 if sum0 > 0.0 else 0.0
            if zfix > 0 and fac > 0.0:
                for k in range(1, km):
                    qc[k] = max(fac * dm[k] / d[k], 0.0)
            q[i, j, :km] = qc


def fv_setup(qvapor, qliquid, qrain, qsnow, qice, qgraupel, q_con, cvm, pkz, pt, cappa, delp, delz, dp1, nx, ny, nz, halo=3):
    """fv_setup (moist_cv.py:175-234, moist_phys) with moist_cv_nwat6_fn (:32-53); compute domain; outputs in place."""
    from .constants import C_ICE, C_LIQ, CV_AIR, CV_VAP, RDG, RDGAS, ZVIR

    s = (slice(halo, halo + nx), slice(halo, halo + ny), slice(0, nz))
    ql = qliquid[s] + qrain[s]
    qs = qice[s] + qsnow[s] + qgraupel[s]
    gz = ql + qs
    cv = (1.0 - (qvapor[s] + gz)) * CV_AIR + qvapor[s] * CV_VAP + ql * C_LIQ + qs * C_ICE
    cvm[s] = cv
    q_con[s] = gz
    d1 = ZVIR * qvapor[s]
    dp1[s] = d1
    cp = RDGAS / (RDGAS + cv / (1.0 + d1))
    cappa[s] = cp
    pkz[s] = np.exp(cp * np.log(RDG * delp[s] * pt[s] * (1.0 + d1) * (1.0 - gz) / delz[s]))


def lagrangian_to_eulerian(a, nx, ny, nz, halo=3):
    """LagrangianToEulerian.__call__ (remapping.py:485-695, do_sat_adj off, kord 9): init_pe (:42-50),
    moist_cv_pt_pressure (:73-151), pn2_pk_delp (:154-166), map_single(pt) / mapn_tracer / fillz / map_single(w, delz),
    undo_delz_adjust_and_copy_peln (:53-68), moist_pkz (moist_cv.py:112-141), pressures_mapu / _mapv (:196-254) with
    map_single(u, v), update_ua + copy_from_below (:257-272), moist_pt_last_step / adjust_divide (:674-695).
    `a` maps the argument names of the reference call to [i, j, k] arrays, updated in place."""
    from .constants import C_ICE, C_LIQ, CV_AIR, CV_VAP, RDG, RDGAS

    names = ["qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke"]
    tr = {n: a["tracers." + n] for n in names}
    pt, delp, delz, peln, pe, pk, pkz = a["pt"], a["delp"], a["delz"], a["peln"], a["pe"], a["pk"], a["pkz"]
    ak, bk = a["ak"], a["bk"]
    ptop, akap, zvir = float(a["ptop"]), float(a["akap"]), float(a["zvir"])
    ci, cj, cjp, cip = slice(halo, halo + nx), slice(halo, halo + ny), slice(halo, halo + ny + 1), slice(halo, halo + nx + 1)
    K = slice(0, nz)

    def moist_cv():
        ql = tr["qliquid"][ci, cj, K] + tr["qrain"][ci, cj, K]
        qs = tr["qice"][ci, cj, K] + tr["qsnow"][ci, cj, K] + tr["qgraupel"][ci, cj, K]
        gz = ql + qs
        qv = tr["qvapor"][ci, cj, K]
        return (1.0 - (qv + gz)) * CV_AIR + qv * CV_VAP + ql * C_LIQ + qs * C_ICE, gz

    pe1 = pe.copy()
    pe2 = np.zeros_like(pe)
    pe2[ci, cjp, 0] = ptop
    pe2[ci, cjp, nz] = pe[ci, cjp, nz]
    cvm, gz = moist_cv()
    a["q_con"][ci, cj, K] = gz
    cp = RDGAS / (RDGAS + cvm / (1.0 + zvir * tr["qvapor"][ci, cj, K]))
    a["cappa"][ci, cj, K] = cp
    pt[ci, cj, K] = pt[ci, cj, K] * np.exp(cp / (1.0 - cp) * np.log(RDG * delp[ci, cj, K] / delz[ci, cj, K] * pt[ci, cj, K]))
    delz[ci, cj, K] = -delz[ci, cj, K] / delp[ci, cj, K]
    psv = pe[ci, cj, nz].copy()
    a["ps"][ci, cj] = psv
    for k in range(1, nz):
        pe2[ci, cj, k] = ak[k] + bk[k] * psv
    dp2 = np.zeros_like(pe)
    dp2[ci, cj, K] = pe2[ci, cj, 1 : nz + 1] - pe2[ci, cj, K]
    delp[ci, cj, K] = dp2[ci, cj, K]
    pn2 = np.zeros_like(pe)
    pn2[ci, cj, nz] = peln[ci, cj, nz]
    pn2[ci, cj, K] = np.log(pe2[ci, cj, K])
    pk[ci, cj, K] = np.exp(akap * pn2[ci, cj, K])
    map_single(pt, peln, pn2, None, 184.0, 1, nx, ny, nz)
    for n in names:
        map_single(tr[n], pe1, pe2, None, 0.0, 0, nx, ny, nz)
    for n in names:
        fillz(tr[n], dp2, nx, ny, nz)
    map_single(a["w"], pe1, pe2, a["wsd"], 0.0, -2, nx, ny, nz)
    map_single(delz, pe1, pe2, None, 0.0, 1, nx, ny, nz)
    delz[ci, cj, K] = -delz[ci, cj, K] * delp[ci, cj, K]
    peln[ci, cj, : nz + 1] = pn2[ci, cj, : nz + 1]
    cvm, gz = moist_cv()
    a["q_con"][ci, cj, K] = gz
    cp = RDGAS / (RDGAS + cvm / (1.0 + zvir * tr["qvapor"][ci, cj, K]))
    a["cappa"][ci, cj, K] = cp
    pkz[ci, cj, K] = np.exp(cp * np.log(RDG * delp[ci, cj, K] / delz[ci, cj, K] * pt[ci, cj, K]))
    # u on (nx, ny+1): pressures averaged across the y interface
    pe0 = np.zeros_like(pe)
    pe3 = np.zeros_like(pe)
    pem = np.roll(pe, 1, axis=1)     # pe[i, j-1]
    pe0[ci, cjp, 0] = pe[ci, cjp, 0]
    pe0[ci, cjp, 1 : nz + 1] = 0.5 * (pem[ci, cjp, 1 : nz + 1] + pe[ci, cjp, 1 : nz + 1])
    for k in range(nz + 1):
        pe3[ci, cjp, k] = ak[k] + 0.5 * bk[k] * (pem[ci, cjp, nz] + pe[ci, cjp, nz])
    map_single(a["u"], pe0, pe3, None, 0.0, -1, nx, ny, nz, j_extra=1)
    pem = np.roll(pe, 1, axis=0)     # pe[i-1, j]
    pe0[cip, cj, 0] = pe[cip, cj, 0]
    pe0[cip, cj, 1 : nz + 1] = 0.5 * (pem[cip, cj, 1 : nz + 1] + pe[cip, cj, 1 : nz + 1])
    pe3[cip, cj, 0] = ak[0]
    for k in range(1, nz + 1):
        pe3[cip, cj, k] = ak[k] + 0.5 * bk[k] * (pem[cip, cj, nz] + pe[cip, cj, nz])
    map_single(a["v"], pe0, pe3, None, 0.0, -1, nx, ny, nz, i_extra=1)
    pe[ci, cj, 1:nz] = pe2[ci, cj, 1:nz]
    if bool(a["last_step"]):
        gzl = tr["qliquid"][ci, cj, K] + tr["qrain"][ci, cj, K] + tr["qice"][ci, cj, K] + tr["qsnow"][ci, cj, K] + tr["qgraupel"][ci, cj, K]
        pt[ci, cj, K] = (pt[ci, cj, K] + 0.0 * pkz[ci, cj, K]) / ((1.0 + zvir * tr["qvapor"][ci, cj, K]) * (1.0 - gzl))
    else:
        pt[ci, cj, K] = pt[ci, cj, K] / pkz[ci, cj, K]
