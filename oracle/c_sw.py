"""Oracle: C-grid shallow-water half step (test infrastructure).

Follows fv3core/pace/fv3core/stencils/d2a2c_vect.py:19-655 (DGrid2AGrid2CGridVectors) and c_sw.py:19-766
(CGridShallowWaterDynamics) statement by statement.  All arrays [i, j, k], k restricted to [0, nz).
"""
import numpy as np

from .indexing import Idx, sl

A1 = 9.0 / 16.0
A2 = -1.0 / 16.0
C1 = -2.0 / 14.0
C2 = 11.0 / 14.0
C3 = 5.0 / 14.0
BIG = 1e30


def _sh(a, di, dj, si, sj):
    """a shifted: values a[i+di, j+dj] for (i, j) in the slices (si, sj)"""
    return a[si.start + di : si.stop + di, sj.start + dj : sj.stop + dj]


def contravariant(v1, v2, cosa, rsin2):
    return (v1 - v2 * cosa) * rsin2


def d2a2c_vect(ix: Idx, g, uc, vc, u, v, ua, va, utc, vtc):
    """d2a2c_vect.py:547-655.  g: dict of 2-D metric arrays.  uc, vc, ua, va, utc, vtc updated in place."""
    nz = ix.nz
    K = slice(0, nz)
    isc, iec, jsc, jec, ied, jed = ix.isc, ix.iec, ix.jsc, ix.jec, ix.ied, ix.jed
    utmp = np.full(u.shape, BIG)
    vtmp = np.full(u.shape, BIG)
    utmp[:, :, nz:] = 0.0
    vtmp[:, :, nz:] = 0.0
    npt = 4
    if npt > ix.nx - 1 or npt > ix.ny - 1:
        npt = 0
    nxx, nyy = iec + 1, jec + 1
    js1 = npt + 2 if ix.south else jsc - 1
    je1 = nyy - npt if ix.north else jec + 1
    is1 = npt + 2 if ix.west else 0
    ie1 = nxx - npt if ix.east else ied
    is2 = npt + 2 if ix.west else isc - 1
    ie2 = nxx - npt if ix.east else iec + 1
    js2 = npt + 2 if ix.south else 0
    je2 = nyy - npt if ix.north else jed
    # lagrange_interpolation_y_p1 (:27-33) and _x_p1 (:36-43)
    si, sj = sl(is1, ie1), sl(js1, je1)
    utmp[si, sj, K] = A2 * (_sh(u, 0, -1, si, sj)[:, :, K] + _sh(u, 0, 2, si, sj)[:, :, K]) + A1 * (u[si, sj, K] + _sh(u, 0, 1, si, sj)[:, :, K])
    si, sj = sl(is2, ie2), sl(js2, je2)
    vtmp[si, sj, K] = A2 * (_sh(v, -1, 0, si, sj)[:, :, K] + _sh(v, 2, 0, si, sj)[:, :, K]) + A1 * (v[si, sj, K] + _sh(v, 1, 0, si, sj)[:, :, K])
    # avg_box (:46-65) over the full domain
    off = -1 if npt == 0 else 3
    mask = np.zeros(u.shape[:2], dtype=bool)
    if ix.south:
        mask[: ied + 1, : jsc + off] = True
    if ix.north:
        mask[: ied + 1, jec - off + 1 : jed + 1] = True
    if ix.west:
        mask[: isc + off, : jed + 1] = True
    if ix.east:
        mask[iec - off + 1 : ied + 1, : jed + 1] = True
    full_i, full_j = sl(0, ied), sl(0, jed)
    ua_avg = 0.5 * (u[full_i, full_j, K] + _sh(u, 0, 1, full_i, full_j)[:, :, K])
    va_avg = 0.5 * (v[full_i, full_j, K] + _sh(v, 1, 0, full_i, full_j)[:, :, K])
    m3 = mask[full_i, full_j][:, :, None]
    utmp[full_i, full_j, K] = np.where(m3, ua_avg, utmp[full_i, full_j, K])
    vtmp[full_i, full_j, K] = np.where(m3, va_avg, vtmp[full_i, full_j, K])
    # contravariant_components (:68-78) on compute + 2
    si, sj = sl(isc - 2, iec + 2), sl(jsc - 2, jec + 2)
    cs, r2 = g["cosa_s"][si, sj, None], g["rsin2"][si, sj, None]
    ua[si, sj, K] = contravariant(utmp[si, sj, K], vtmp[si, sj, K], cs, r2)
    va[si, sj, K] = contravariant(vtmp[si, sj, K], utmp[si, sj, K], cs, r2)
    # fill_corners_x (:81-88): utmp 3 cells, ua 2 cells from vtmp / va (corners.py:130-232)
    mult = dict(sw=-1.0, se=1.0, ne=-1.0, nw=1.0)
    _fill_corners_x(ix, utmp, vtmp, 3, mult, K)
    _fill_corners_x(ix, ua, va, 2, mult, K)
    # ut_main (:91-101)
    ifirst = isc + 2 if ix.west else isc - 1
    ilast = iec - 1 if ix.east else iec + 2
    si, sj = sl(ifirst, ilast), sl(jsc - 1, jec + 1)
    uc[si, sj, K] = A2 * (_sh(utmp, -2, 0, si, sj)[:, :, K] + _sh(utmp, 1, 0, si, sj)[:, :, K]) + A1 * (_sh(utmp, -1, 0, si, sj)[:, :, K] + utmp[si, sj, K])
    utc[si, sj, K] = contravariant(uc[si, sj, K], v[si, sj, K], g["cosa_u"][si, sj, None], g["rsin_u"][si, sj, None])
    # east_west_edges (:104-154)
    sj = sl(jsc - 1, jec + 1)

    def x_edge(i0):
        # i0 = i_start (west) or i_end + 1 (east): columns i0-1 (cubic), i0 (edge interp), i0+1 (cubic rev)
        i = i0 - 1
        uc[i, sj, K] = C1 * utmp[i - 2, sj, K] + C2 * utmp[i - 1, sj, K] + C3 * utmp[i, sj, K]
        i = i0
        dxa = g["dxa"]
        t1 = dxa[i - 2, sj] + dxa[i - 1, sj]
        t2 = dxa[i, sj] + dxa[i + 1, sj]
        n1 = (t1 + dxa[i - 1, sj])[:, None] * ua[i - 1, sj, K] - dxa[i - 1, sj][:, None] * ua[i - 2, sj, K]
        n2 = (t1 + dxa[i, sj])[:, None] * ua[i, sj, K] - dxa[i, sj][:, None] * ua[i + 1, sj, K]
        e = 0.5 * (n1 / t1[:, None] + n2 / t2[:, None])
        utc[i, sj, K] = e
        uc[i, sj, K] = np.where(e > 0, e * g["sin_sg3"][i - 1, sj][:, None], e * g["sin_sg1"][i, sj][:, None])
        i = i0 + 1
        uc[i, sj, K] = C1 * utmp[i + 1, sj, K] + C2 * utmp[i, sj, K] + C3 * utmp[i - 1, sj, K]
        for i in (i0 - 1, i0 + 1):
            utc[i, sj, K] = contravariant(uc[i, sj, K], v[i, sj, K], g["cosa_u"][i, sj][:, None], g["rsin_u"][i, sj][:, None])

    if ix.west:
        x_edge(isc)
    if ix.east:
        x_edge(iec + 1)
    # fill_corners_y (:157-164)
    _fill_corners_y(ix, vtmp, utmp, 3, mult, K)
    _fill_corners_y(ix, va, ua, 2, mult, K)
    # north_south_edges (:167-215) + vt_main (:218-228)
    si, sj = sl(isc - 1, iec + 1), sl(jsc - 1, jec + 2)
    vc[si, sj, K] = A2 * (_sh(vtmp, 0, -2, si, sj)[:, :, K] + _sh(vtmp, 0, 1, si, sj)[:, :, K]) + A1 * (_sh(vtmp, 0, -1, si, sj)[:, :, K] + vtmp[si, sj, K])
    vtc[si, sj, K] = contravariant(vc[si, sj, K], u[si, sj, K], g["cosa_v"][si, sj, None], g["rsin_v"][si, sj, None])

    def y_edge(j0):
        j = j0 - 1
        vc[si, j, K] = C1 * vtmp[si, j - 2, K] + C2 * vtmp[si, j - 1, K] + C3 * vtmp[si, j, K]
        j = j0
        dya = g["dya"]
        t1 = dya[si, j - 2] + dya[si, j - 1]
        t2 = dya[si, j] + dya[si, j + 1]
        n1 = (t1 + dya[si, j - 1])[:, None] * va[si, j - 1, K] - dya[si, j - 1][:, None] * va[si, j - 2, K]
        n2 = (t1 + dya[si, j])[:, None] * va[si, j, K] - dya[si, j][:, None] * va[si, j + 1, K]
        e = 0.5 * (n1 / t1[:, None] + n2 / t2[:, None])
        vtc[si, j, K] = e
        vc[si, j, K] = np.where(e > 0, e * g["sin_sg4"][si, j - 1][:, None], e * g["sin_sg2"][si, j][:, None])
        j = j0 + 1
        vc[si, j, K] = C1 * vtmp[si, j + 1, K] + C2 * vtmp[si, j, K] + C3 * vtmp[si, j - 1, K]
        for j in (j0 - 1, j0 + 1):
            vtc[si, j, K] = contravariant(vc[si, j, K], u[si, j, K], g["cosa_v"][si, j][:, None], g["rsin_v"][si, j][:, None])

    if ix.south:
        y_edge(jsc)
    if ix.north:
        y_edge(jec + 1)
    return utmp, vtmp


def _fill_corners_x(ix, q, qc, ncells, mult, K):
    """corners.py:130-232 fill_corners_{2,3}cells_mult_x"""
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    for d in range(1, ncells + 1):
        if ix.south and ix.west:
            q[isc - d, jsc - 1, K] = mult["sw"] * qc[isc - 1, jsc - 1 + d, K]
        if ix.south and ix.east:
            q[iec + d, jsc - 1, K] = mult["se"] * qc[iec + 1, jsc - 1 + d, K]
        if ix.north and ix.west:
            q[isc - d, jec + 1, K] = mult["nw"] * qc[isc - 1, jec + 1 - d, K]
        if ix.north and ix.east:
            q[iec + d, jec + 1, K] = mult["ne"] * qc[iec + 1, jec + 1 - d, K]


def _fill_corners_y(ix, q, qc, ncells, mult, K):
    """corners.py:235-304 fill_corners_{2,3}cells_mult_y"""
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    for d in range(1, ncells + 1):
        if ix.south and ix.west:
            q[isc - 1, jsc - d, K] = mult["sw"] * qc[isc - 1 + d, jsc - 1, K]
        if ix.south and ix.east:
            q[iec + 1, jsc - d, K] = mult["se"] * qc[iec + 1 - d, jsc - 1, K]
        if ix.north and ix.west:
            q[isc - 1, jec + d, K] = mult["nw"] * qc[isc - 1 + d, jec + 1, K]
        if ix.north and ix.east:
            q[iec + 1, jec + d, K] = mult["ne"] * qc[iec + 1 - d, jec + 1, K]


def divergence_corner(ix: Idx, g, u, v, ua, va, divg_d):
    """c_sw.py:31-154 on the (nx+1) x (ny+1) corner points of the compute domain."""
    K = slice(0, ix.nz)
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    # uf, vf on an extended domain (one extra row/column towards -i / -j)
    si, sj = sl(isc - 1, iec + 1), sl(jsc - 1, jec + 1)

    def e2(name, di=0, dj=0):
        return _sh(g[name], di, dj, si, sj)[:, :, None]

    uf = ((u[si, sj, K] - 0.25 * (_sh(va, 0, -1, si, sj)[:, :, K] + va[si, sj, K]) * (e2("cos_sg4", 0, -1) + e2("cos_sg2")))
          * e2("dyc") * 0.5 * (e2("sin_sg4", 0, -1) + e2("sin_sg2")))
    vf = ((v[si, sj, K] - 0.25 * (_sh(ua, -1, 0, si, sj)[:, :, K] + ua[si, sj, K]) * (e2("cos_sg3", -1, 0) + e2("cos_sg1")))
          * e2("dxc") * 0.5 * (e2("sin_sg3", -1, 0) + e2("sin_sg1")))
    uf_e = u[si, sj, K] * e2("dyc") * 0.5 * (e2("sin_sg4", 0, -1) + e2("sin_sg2"))
    vf_e = v[si, sj, K] * e2("dxc") * 0.5 * (e2("sin_sg3", -1, 0) + e2("sin_sg1"))
    # local index helpers: array index a <-> storage index a + isc - 1
    oi, oj = isc - 1, jsc - 1
    if ix.west:
        vf[isc - oi, :, :] = vf_e[isc - oi, :, :]
    if ix.east:
        vf[iec + 1 - oi, :, :] = vf_e[iec + 1 - oi, :, :]
    if ix.south:
        uf[:, jsc - oj, :] = uf_e[:, jsc - oj, :]
    if ix.north:
        uf[:, jec + 1 - oj, :] = uf_e[:, jec + 1 - oj, :]
    ci, cj = sl(isc, iec + 1), sl(jsc, jec + 1)
    rc = g["rarea_c"][ci, cj, None]
    vf1 = vf[1:, :-1]
    vf0 = vf[1:, 1:]
    uf1 = uf[:-1, 1:]
    uf0 = uf[1:, 1:]
    out = (vf1 - vf0 + uf1 - uf0) * rc
    for (ci_, on_i) in ((0, ix.west), (ix.nx, ix.east)):
        if ix.south and on_i:
            out[ci_, 0] = (-vf0[ci_, 0] + uf1[ci_, 0] - uf0[ci_, 0]) * rc[ci_, 0]
        if ix.north and on_i:
            out[ci_, ix.ny] = (vf1[ci_, ix.ny] + uf1[ci_, ix.ny] - uf0[ci_, ix.ny]) * rc[ci_, ix.ny]
    divg_d[ci, cj, K] = out


def _fill2_x(ix, q, K):
    _fill_corners_x(ix, q, q, 2, dict(sw=1.0, se=1.0, nw=1.0, ne=1.0), K)


def _fill2_y(ix, q, K):
    _fill_corners_y(ix, q, q, 2, dict(sw=1.0, se=1.0, nw=1.0, ne=1.0), K)


def c_sw(ix: Idx, g, delp, pt, u, v, w, uc, vc, ua, va, ut, vt, divgd, omga, dt2, nord=3):
    """CGridShallowWaterDynamics.__call__ (c_sw.py:607-766).  Returns (delpc, ptc); everything else in place."""
    nz = ix.nz
    K = slice(0, nz)
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    delpc = np.zeros_like(delp)
    ptc = np.zeros_like(delp)
    d2a2c_vect(ix, g, uc, vc, u, v, ua, va, ut, vt)
    if nord > 0:
        divergence_corner(ix, g, u, v, ua, va, divgd)
    # geoadjust_ut / vt (:156-199)
    si, sj = sl(isc - 1, iec + 2), sl(jsc - 1, jec + 1)
    a = ut[si, sj, K]
    ut[si, sj, K] = np.where(a > 0, dt2 * a * g["dy"][si, sj, None] * _sh(g["sin_sg3"], -1, 0, si, sj)[:, :, None],
                             dt2 * a * g["dy"][si, sj, None] * g["sin_sg1"][si, sj, None])
    si, sj = sl(isc - 1, iec + 1), sl(jsc - 1, jec + 2)
    a = vt[si, sj, K]
    vt[si, sj, K] = np.where(a > 0, dt2 * a * g["dx"][si, sj, None] * _sh(g["sin_sg4"], 0, -1, si, sj)[:, :, None],
                             dt2 * a * g["dx"][si, sj, None] * g["sin_sg2"][si, sj, None])
    # fill_corners x, x-fluxes (:229-258)
    for q in (delp, pt, w):
        _fill2_x(ix, q, K)
    si, sj = sl(isc - 1, iec + 2), sl(jsc - 1, jec + 1)
    up = ut[si, sj, K] > 0.0
    fx1 = np.where(up, _sh(delp, -1, 0, si, sj)[:, :, K], delp[si, sj, K])
    fx = np.where(up, _sh(pt, -1, 0, si, sj)[:, :, K], pt[si, sj, K])
    fx2 = np.where(up, _sh(w, -1, 0, si, sj)[:, :, K], w[si, sj, K])
    fx1 = ut[si, sj, K] * fx1
    fx = fx1 * fx
    fx2 = fx1 * fx2
    for q in (delp, pt, w):
        _fill2_y(ix, q, K)
    # transportdelp_update_vorticity_and_kineticenergy (:261-364)
    si, sj = sl(isc - 1, iec + 1), sl(jsc - 1, jec + 2)
    up = vt[si, sj, K] > 0.0
    fy1 = np.where(up, _sh(delp, 0, -1, si, sj)[:, :, K], delp[si, sj, K])
    fy = np.where(up, _sh(pt, 0, -1, si, sj)[:, :, K], pt[si, sj, K])
    fy2 = np.where(up, _sh(w, 0, -1, si, sj)[:, :, K], w[si, sj, K])
    fy1 = vt[si, sj, K] * fy1
    fy = fy1 * fy
    fy2 = fy1 * fy2
    si, sj = sl(isc - 1, iec + 1), sl(jsc - 1, jec + 1)
    ra = g["rarea"][si, sj, None]
    delpc[si, sj, K] = delp[si, sj, K] + (fx1[:-1] - fx1[1:] + fy1[:, :-1] - fy1[:, 1:]) * ra
    ptc[si, sj, K] = (pt[si, sj, K] * delp[si, sj, K] + (fx[:-1] - fx[1:] + fy[:, :-1] - fy[:, 1:]) * ra) / delpc[si, sj, K]
    omga[si, sj, K] = (w[si, sj, K] * delp[si, sj, K] + (fx2[:-1] - fx2[1:] + fy2[:, :-1] - fy2[:, 1:]) * ra) / delpc[si, sj, K]
    uap, vap = ua[si, sj, K] > 0.0, va[si, sj, K] > 0.0
    ke = np.where(uap, uc[si, sj, K], _sh(uc, 1, 0, si, sj)[:, :, K])
    vort = np.where(vap, vc[si, sj, K], _sh(vc, 0, 1, si, sj)[:, :, K])
    oi, oj = isc - 1, jsc - 1

    def g2(name):
        return g[name][si, sj, None]

    u_n = _sh(u, 0, 1, si, sj)[:, :, K]
    v_e = _sh(v, 1, 0, si, sj)[:, :, K]
    uu, vv = u[si, sj, K], v[si, sj, K]
    for (flag, j) in ((ix.south, jsc - 1), (ix.north, jec)):
        if flag:
            jj = j - oj
            vort[:, jj] = np.where(~vap[:, jj], vort[:, jj] * g2("sin_sg4")[:, jj] + u_n[:, jj] * g2("cos_sg4")[:, jj], vort[:, jj])
    for (flag, j) in ((ix.south, jsc), (ix.north, jec + 1)):
        if flag:
            jj = j - oj
            vort[:, jj] = np.where(vap[:, jj], vort[:, jj] * g2("sin_sg2")[:, jj] + uu[:, jj] * g2("cos_sg2")[:, jj], vort[:, jj])
    for (flag, i) in ((ix.east, iec), (ix.west, isc - 1)):
        if flag:
            ii = i - oi
            ke[ii] = np.where(~uap[ii], ke[ii] * g2("sin_sg3")[ii] + v_e[ii] * g2("cos_sg3")[ii], ke[ii])
    for (flag, i) in ((ix.east, iec + 1), (ix.west, isc)):
        if flag:
            ii = i - oi
            ke[ii] = np.where(uap[ii], ke[ii] * g2("sin_sg1")[ii] + vv[ii] * g2("cos_sg1")[ii], ke[ii])
    ke = 0.5 * dt2 * (ua[si, sj, K] * ke + va[si, sj, K] * vort)
    ke_full = np.zeros_like(delp)
    ke_full[si, sj, K] = ke
    # circulation_cgrid (:367-397) + absolute_vorticity (:400-408)
    si, sj = sl(isc, iec + 1), sl(jsc, jec + 1)
    fx_ = g["dxc"][si, sj, None] * uc[si, sj, K]
    fy_ = g["dyc"][si, sj, None] * vc[si, sj, K]
    fx1_ = _sh(g["dxc"], 0, -1, si, sj)[:, :, None] * _sh(uc, 0, -1, si, sj)[:, :, K]
    fy1_ = _sh(g["dyc"], -1, 0, si, sj)[:, :, None] * _sh(vc, -1, 0, si, sj)[:, :, K]
    vort_c = fx1_ - fx_ - fy1_ + fy_
    for (fi, ii) in ((ix.west, 0),):
        for (fj, jj) in ((ix.south, 0), (ix.north, ix.ny)):
            if fi and fj:
                vort_c[ii, jj] = fx1_[ii, jj] - fx_[ii, jj] + fy_[ii, jj]
    for (fi, ii) in ((ix.east, ix.nx),):
        for (fj, jj) in ((ix.south, 0), (ix.north, ix.ny)):
            if fi and fj:
                vort_c[ii, jj] = fx1_[ii, jj] - fx_[ii, jj] - fy1_[ii, jj]
    vort_c = g["fC"][si, sj, None] + g["rarea_c"][si, sj, None] * vort_c
    vfull = np.zeros_like(delp)
    vfull[si, sj, K] = vort_c
    # update_y_velocity (:455-480): vc on [isc..iec] x [jsc..jec+1]
    si, sj = sl(isc, iec), sl(jsc, jec + 1)
    tmp = dt2 * (u[si, sj, K] - vc[si, sj, K] * g["cosa_v"][si, sj, None]) / g["sina_v"][si, sj, None]
    if ix.south:
        tmp[:, 0] = dt2 * u[si, jsc, K]
    if ix.north:
        tmp[:, ix.ny] = dt2 * u[si, jec + 1, K]
    flux = np.where(tmp > 0.0, vfull[si, sj, K], _sh(vfull, 1, 0, si, sj)[:, :, K])
    vc[si, sj, K] = vc[si, sj, K] - tmp * flux + g["rdyc"][si, sj, None] * (_sh(ke_full, 0, -1, si, sj)[:, :, K] - ke_full[si, sj, K])
    # update_x_velocity (:411-452): uc on [isc..iec+1] x [jsc..jec]
    si, sj = sl(isc, iec + 1), sl(jsc, jec)
    tmp = dt2 * (v[si, sj, K] - uc[si, sj, K] * g["cosa_u"][si, sj, None]) / g["sina_u"][si, sj, None]
    if ix.west:
        tmp[0] = dt2 * v[isc, sj, K]
    if ix.east:
        tmp[ix.nx] = dt2 * v[iec + 1, sj, K]
    flux = np.where(tmp > 0.0, vfull[si, sj, K], _sh(vfull, 0, 1, si, sj)[:, :, K])
    uc[si, sj, K] = uc[si, sj, K] + tmp * flux + g["rdxc"][si, sj, None] * (_sh(ke_full, -1, 0, si, sj)[:, :, K] - ke_full[si, sj, K])
    return delpc, ptc
