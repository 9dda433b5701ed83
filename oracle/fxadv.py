"""Oracle: FiniteVolumeFluxPrep (fv3core/pace/fv3core/stencils/fxadv.py:9-661) — test infrastructure.

Statement-by-statement restatement including the reference's save/restore idiom (utmp/vtmp) and the aliased
uc_contra_copy / vc_contra_copy arguments (fxadv.py:625-644).
"""
import numpy as np

from .c_sw import _sh, contravariant
from .indexing import Idx, sl


def fv_prep(ix: Idx, g, uc, vc, crx, cry, xfx, yfx, uc_contra, vc_contra, dt):
    nz = ix.nz
    K = slice(0, nz)
    isc, iec, jsc, jec, ied, jed = ix.isc, ix.iec, ix.jsc, ix.jec, ix.ied, ix.jed
    FI, FJ = sl(0, ied), sl(0, jed)

    def m2(name, si, sj, di=0, dj=0):
        return _sh(g[name], di, dj, si, sj)[:, :, None]

    # main_uc_vc_contra (:9-40)
    utmp = uc_contra.copy()
    si, sj = sl(isc - 1, iec + 2), FJ
    v = 0.25 * (_sh(vc, -1, 0, si, sj)[:, :, K] + vc[si, sj, K] + _sh(vc, -1, 1, si, sj)[:, :, K] + _sh(vc, 0, 1, si, sj)[:, :, K])
    uc_contra[si, sj, K] = contravariant(uc[si, sj, K], v, m2("cosa_u", si, sj), m2("rsin_u", si, sj))
    for flag, rows in ((ix.south, sl(jsc - 1, jsc)), (ix.north, sl(jec, jec + 1))):
        if flag:
            uc_contra[FI, rows, K] = utmp[FI, rows, K]
    si, sj = FI, sl(jsc - 1, jec + 2)
    u = 0.25 * (_sh(uc, 0, -1, si, sj)[:, :, K] + _sh(uc, 1, -1, si, sj)[:, :, K] + uc[si, sj, K] + _sh(uc, 1, 0, si, sj)[:, :, K])
    vc_contra[si, sj, K] = contravariant(vc[si, sj, K], u, m2("cosa_v", si, sj), m2("rsin_v", si, sj))
    if ix.west or ix.east or ix.south or ix.north:
        # uc_contra_y_edge (:43-58)
        for flag, i in ((ix.west, isc), (ix.east, iec + 1)):
            if flag:
                a = uc[i, FJ, K]
                uc_contra[i, FJ, K] = np.where(a > 0, a / g["sin_sg3"][i - 1, FJ, None], a / g["sin_sg1"][i, FJ, None])
        # vc_contra_y_edge (:61-90)
        vtmp = vc_contra.copy()
        for flag, cols in ((ix.west, sl(isc - 1, isc)), (ix.east, sl(iec, iec + 1))):
            if flag:
                si, sj = cols, sl(jsc, jec + 1)
                uco = 0.25 * (_sh(uc_contra, 0, -1, si, sj)[:, :, K] + _sh(uc_contra, 1, -1, si, sj)[:, :, K]
                              + uc_contra[si, sj, K] + _sh(uc_contra, 1, 0, si, sj)[:, :, K])
                vc_contra[si, sj, K] = contravariant(vc[si, sj, K], uco, m2("cosa_v", si, sj), 1.0)
                for f2, rows in ((ix.south, sl(jsc, jsc + 1)), (ix.north, sl(jec, jec + 1))):
                    if f2:
                        vc_contra[si, rows, K] = vtmp[si, rows, K]
        # vc_contra_x_edge (:93-104)
        for flag, j in ((ix.south, jsc), (ix.north, jec + 1)):
            if flag:
                a = vc[FI, j, K]
                vc_contra[FI, j, K] = np.where(a > 0, a / g["sin_sg4"][FI, j - 1, None], a / g["sin_sg2"][FI, j, None])
        # uc_contra_x_edge (:107-133)
        utmp = uc_contra.copy()
        for flag, rows in ((ix.south, sl(jsc - 1, jsc)), (ix.north, sl(jec, jec + 1))):
            if flag:
                si, sj = sl(isc, iec + 1), rows
                vco = 0.25 * (_sh(vc_contra, -1, 0, si, sj)[:, :, K] + vc_contra[si, sj, K]
                              + _sh(vc_contra, -1, 1, si, sj)[:, :, K] + _sh(vc_contra, 0, 1, si, sj)[:, :, K])
                uc_contra[si, sj, K] = contravariant(uc[si, sj, K], vco, m2("cosa_u", si, sj), 1.0)
                for f2, cols in ((ix.west, sl(isc, isc + 1)), (ix.east, sl(iec, iec + 1))):
                    if f2:
                        uc_contra[cols, sj, K] = utmp[cols, sj, K]
        _uc_contra_corners(ix, g, uc, vc, uc_contra, vc_contra, K)
        _vc_contra_corners(ix, g, uc, vc, uc_contra, vc_contra, K)
    # fxadv_fluxes_stencil (:355-390)
    si, sj = sl(isc, iec + 1), FJ
    a = uc_contra[si, sj, K]
    crx[si, sj, K] = np.where(a > 0, dt * a * m2("rdxa", si, sj, -1, 0), dt * a * m2("rdxa", si, sj))
    xfx[si, sj, K] = np.where(a > 0, m2("dy", si, sj) * dt * a * m2("sin_sg3", si, sj, -1, 0), m2("dy", si, sj) * dt * a * m2("sin_sg1", si, sj))
    si, sj = FI, sl(jsc, jec + 1)
    a = vc_contra[si, sj, K]
    cry[si, sj, K] = np.where(a > 0, dt * a * m2("rdya", si, sj, 0, -1), dt * a * m2("rdya", si, sj))
    yfx[si, sj, K] = np.where(a > 0, m2("dx", si, sj) * dt * a * m2("sin_sg4", si, sj, 0, -1), m2("dx", si, sj) * dt * a * m2("sin_sg2", si, sj))


def _uc_contra_corners(ix, g, uc, vc, ucc, vcc, K):
    """fxadv.py:136-243"""
    cu, cv = g["cosa_u"], g["cosa_v"]
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec

    def at(a, i, j):
        return a[i, j, K]

    for fj, (ja, jb) in ((ix.south, (jsc - 1, jsc)), (ix.north, (jec, jec + 1))):
        if not fj:
            continue
        if ix.west:
            i, j = isc + 1, ja
            damp = 1.0 / (1.0 - 0.0625 * cu[i, j] * cv[i - 1, j])
            new_a = (at(uc, i, j) - 0.25 * cu[i, j] * (at(vcc, i - 1, j + 1) + at(vcc, i, j + 1) + at(vcc, i, j) + at(vc, i - 1, j)
                     - 0.25 * cv[i - 1, j] * (at(ucc, i - 1, j) + at(ucc, i - 1, j - 1) + at(ucc, i, j - 1)))) * damp
            ucc[i, j, K] = new_a
            j = jb
            damp = 1.0 / (1.0 - 0.0625 * cu[i, j] * cv[i - 1, j + 1])
            ucc[i, j, K] = (at(uc, i, j) - 0.25 * cu[i, j] * (at(vcc, i - 1, j) + at(vcc, i, j) + at(vcc, i, j + 1) + at(vc, i - 1, j + 1)
                            - 0.25 * cv[i - 1, j + 1] * (at(ucc, i - 1, j) + at(ucc, i - 1, j + 1) + at(ucc, i, j + 1)))) * damp
    for fj, (ja, jb) in ((ix.south, (jsc - 1, jsc)), (ix.north, (jec, jec + 1))):
        if not fj:
            continue
        if ix.east:
            i, j = iec, ja
            damp = 1.0 / (1.0 - 0.0625 * cu[i, j] * cv[i, j])
            ucc[i, j, K] = (at(uc, i, j) - 0.25 * cu[i, j] * (at(vcc, i, j + 1) + at(vcc, i - 1, j + 1) + at(vcc, i - 1, j) + at(vc, i, j)
                            - 0.25 * cv[i, j] * (at(ucc, i + 1, j) + at(ucc, i + 1, j - 1) + at(ucc, i, j - 1)))) * damp
            j = jb
            damp = 1.0 / (1.0 - 0.0625 * cu[i, j] * cv[i, j + 1])
            ucc[i, j, K] = (at(uc, i, j) - 0.25 * cu[i, j] * (at(vcc, i, j) + at(vcc, i - 1, j) + at(vcc, i - 1, j + 1) + at(vc, i, j + 1)
                            - 0.25 * cv[i, j + 1] * (at(ucc, i + 1, j) + at(ucc, i + 1, j + 1) + at(ucc, i, j + 1)))) * damp


def _vc_contra_corners(ix, g, uc, vc, ut, vcc, K):
    """fxadv.py:246-352 (ut = uc_contra)"""
    cu, cv = g["cosa_u"], g["cosa_v"]
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec

    def at(a, i, j):
        return a[i, j, K]

    if ix.south:
        j = jsc + 1
        for fi, i in ((ix.west, isc - 1), (ix.east, iec)):
            if fi:
                damp = 1.0 / (1.0 - 0.0625 * cu[i, j - 1] * cv[i, j])
                vcc[i, j, K] = (at(vc, i, j) - 0.25 * cv[i, j] * (at(ut, i + 1, j - 1) + at(ut, i + 1, j) + at(ut, i, j) + at(uc, i, j - 1)
                                - 0.25 * cu[i, j - 1] * (at(vcc, i, j - 1) + at(vcc, i - 1, j - 1) + at(vcc, i - 1, j)))) * damp
        for fi, i in ((ix.west, isc), (ix.east, iec + 1)):
            if fi:
                damp = 1.0 / (1.0 - 0.0625 * cu[i + 1, j - 1] * cv[i, j])
                vcc[i, j, K] = (at(vc, i, j) - 0.25 * cv[i, j] * (at(ut, i, j - 1) + at(ut, i, j) + at(ut, i + 1, j) + at(uc, i + 1, j - 1)
                                - 0.25 * cu[i + 1, j - 1] * (at(vcc, i, j - 1) + at(vcc, i + 1, j - 1) + at(vcc, i + 1, j)))) * damp
    if ix.north:
        j = jec
        for fi, i in ((ix.east, iec + 1), (ix.west, isc)):
            if fi:
                damp = 1.0 / (1.0 - 0.0625 * cu[i + 1, j] * cv[i, j])
                vcc[i, j, K] = (at(vc, i, j) - 0.25 * cv[i, j] * (at(ut, i, j) + at(ut, i, j - 1) + at(ut, i + 1, j - 1) + at(uc, i + 1, j)
                                - 0.25 * cu[i + 1, j] * (at(vcc, i, j + 1) + at(vcc, i + 1, j + 1) + at(vcc, i + 1, j)))) * damp
        for fi, i in ((ix.east, iec), (ix.west, isc - 1)):
            if fi:
                damp = 1.0 / (1.0 - 0.0625 * cu[i, j] * cv[i, j])
                vcc[i, j, K] = (at(vc, i, j) - 0.25 * cv[i, j] * (at(ut, i + 1, j) + at(ut, i + 1, j - 1) + at(ut, i, j - 1) + at(uc, i, j)
                                - 0.25 * cu[i, j] * (at(vcc, i, j + 1) + at(vcc, i - 1, j + 1) + at(vcc, i - 1, j)))) * damp
