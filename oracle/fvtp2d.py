"""Oracle: 2-D finite-volume transport, PPM sweeps, corner copies and del-n fluxes — test infrastructure.

Follows fv3core/pace/fv3core/stencils/fvtp2d.py:15-346, xppm.py:22-353 (yppm.py is its mirror image), ppm.py:8-36,
delnflux.py:18-1261 and stencils/pace/stencils/corners.py:307-425 of the reference.
"""
import numpy as np

from .indexing import Idx, sl

C1, C2, C3 = -2.0 / 14.0, 11.0 / 14.0, 5.0 / 14.0
P1, P2 = 7.0 / 12.0, -1.0 / 12.0
S11, S14, S15 = 11.0 / 14.0, 4.0 / 7.0, 3.0 / 14.0


def rsign(a, b):
    """basic_operations.sign (basic_operations.py:32-39)"""
    return np.where(b > 0, np.abs(a), -np.abs(a))


# ---------------------------------------------------------------------------------------------
# corner copies (corners.py:307-425): in place, 3x3 blocks, only at cube corners

def copy_corners_x(ix: Idx, q, K):
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    for a in (1, 2, 3):
        for b in (1, 2, 3):
            if ix.west and ix.south:
                q[isc - a, jsc - b, K] = q[isc - b, jsc + a - 1, K]
            if ix.east and ix.south:
                q[iec + a, jsc - b, K] = q[iec + b, jsc + a - 1, K]
            if ix.west and ix.north:
                q[isc - a, jec + b, K] = q[isc - b, jec - a + 1, K]
            if ix.east and ix.north:
                q[iec + a, jec + b, K] = q[iec + b, jec - a + 1, K]


def copy_corners_y(ix: Idx, q, K):
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    for a in (1, 2, 3):
        for b in (1, 2, 3):
            if ix.west and ix.south:
                q[isc - a, jsc - b, K] = q[isc + b - 1, jsc - a, K]
            if ix.east and ix.south:
                q[iec + a, jsc - b, K] = q[iec - b + 1, jsc - a, K]
            if ix.west and ix.north:
                q[isc - a, jec + b, K] = q[isc + b - 1, jec + a, K]
            if ix.east and ix.north:
                q[iec + a, jec + b, K] = q[iec - b + 1, jec + a, K]


# ---------------------------------------------------------------------------------------------
# PPM sweep along axis 0 (xppm.py); for y pass transposed arrays

def _edge01(q, dx, i, which, minmax):
    """xt_dxa_edge_0 (which=0) / xt_dxa_edge_1 (which=1) at cell i (xppm.py:105-145); dx[i] broadcast [n_j, 1]"""
    d = lambda n: dx[n][:, None]  # noqa: E731
    if which == 0:
        xt = 0.5 * (((2.0 * d(i) + d(i - 1)) * q[i] - d(i) * q[i - 1]) / (d(i - 1) + d(i))
                    + ((2.0 * d(i + 1) + d(i + 2)) * q[i + 1] - d(i + 1) * q[i + 2]) / (d(i + 1) + d(i + 2)))
        lo, hi = i - 1, i + 2
    else:
        xt = 0.5 * (((2.0 * d(i - 1) + d(i - 2)) * q[i - 1] - d(i - 1) * q[i - 2]) / (d(i - 2) + d(i - 1))
                    + ((2.0 * d(i) + d(i + 1)) * q[i] - d(i) * q[i + 1]) / (d(i) + d(i + 1)))
        lo, hi = i - 2, i + 1
    if minmax:
        mn = np.minimum.reduce([q[n] for n in range(lo, hi + 1)])
        mx = np.maximum.reduce([q[n] for n in range(lo, hi + 1)])
        xt = np.minimum(np.maximum(xt, mn), mx)
    return xt


def ppm_flux(q, c, dx, ord_, lo, hi, start, end, i0, i1, minmax=True):
    """Value advected through interfaces i0..i1 (inclusive) of a sweep along axis 0.

    q [n, nj, nk], c (courant) [n, nj, nk], dx [n, nj]; lo/hi: on the low/high tile edge; start/end: first/last
    compute index.  compute_x_flux (xppm.py:249-266).
    """
    mord = abs(ord_)
    n = q.shape[0]
    cells = np.arange(i0 - 1, i1 + 1)          # cells whose bl/br are needed
    if mord < 8:
        ifc = np.arange(i0 - 1, i1 + 2)        # interfaces whose al is needed
        al = P1 * (q[ifc - 1] + q[ifc]) + P2 * (q[ifc - 2] + q[ifc + 1])     # compute_al (:148-181)
        d = lambda m_: dx[m_][:, None]  # noqa: E731
        for flag, base in ((lo, start), (hi, end + 1)):
            if not flag:
                continue
            for i, kind in ((base - 1, 0), (base, 1), (base + 1, 2)):
                if i < ifc[0] or i > ifc[-1]:
                    continue
                p = i - ifc[0]
                if kind == 0:
                    al[p] = C1 * q[i - 2] + C2 * q[i - 1] + C3 * q[i]
                elif kind == 1:
                    al[p] = 0.5 * (((2.0 * d(i - 1) + d(i - 2)) * q[i - 1] - d(i - 1) * q[i - 2]) / (d(i - 2) + d(i - 1))
                                   + ((2.0 * d(i) + d(i + 1)) * q[i] - d(i) * q[i + 1]) / (d(i) + d(i + 1)))
                else:
                    al[p] = C3 * q[i - 1] + C2 * q[i] + C1 * q[i + 1]
        bl = al[:-1] - q[cells]
        br = al[1:] - q[cells]
        b0 = bl + br
        smt5 = (bl * br < 0) if mord == 5 else ((3.0 * np.abs(b0)) < np.abs(bl - br))
        mask = np.where(smt5[:-1] | smt5[1:], 1.0, 0.0)
    else:
        dmi = np.arange(i0 - 3, i1 + 3)
        dmi = dmi[(dmi >= 1) & (dmi <= n - 2)]
        dm_full = np.zeros((n,) + q.shape[1:])
        xt = 0.25 * (q[dmi + 1] - q[dmi - 1])                               # dm_iord8plus (:82-89)
        dqr = np.maximum(np.maximum(q[dmi], q[dmi - 1]), q[dmi + 1]) - q[dmi]
        dql = q[dmi] - np.minimum(np.minimum(q[dmi], q[dmi - 1]), q[dmi + 1])
        dm_full[dmi] = rsign(np.minimum(np.minimum(np.abs(xt), dqr), dql), xt)
        ifc = np.arange(max(i0 - 1, 2), min(i1 + 2, n - 2) + 1)
        al_full = np.zeros_like(dm_full)
        al_full[ifc] = 0.5 * (q[ifc - 1] + q[ifc]) + 1.0 / 3.0 * (dm_full[ifc - 1] - dm_full[ifc])   # :92-94
        xt = 2.0 * dm_full[cells]
        bl = -1.0 * rsign(np.minimum(np.abs(xt), np.abs(al_full[cells] - q[cells])), xt)             # :97-102
        br = rsign(np.minimum(np.abs(xt), np.abs(al_full[cells + 1] - q[cells])), xt)
        for flag, base, kinds in ((lo, start, (1, 2, 3)), (hi, end, (4, 5, 6))):                    # bl_br_edges
            if not flag:
                continue
            off = -1 if kinds[0] == 1 else -1
            for i, kind in zip((base + off, base + off + 1, base + off + 2), kinds):
                if i < cells[0] or i > cells[-1]:
                    continue

                def dm_at(cc):
                    x = 0.25 * (q[cc + 1] - q[cc - 1])
                    r = np.maximum(np.maximum(q[cc], q[cc - 1]), q[cc + 1]) - q[cc]
                    l_ = q[cc] - np.minimum(np.minimum(q[cc], q[cc - 1]), q[cc + 1])
                    return rsign(np.minimum(np.minimum(np.abs(x), r), l_), x)

                if kind == 1:
                    xbl = S14 * dm_at(i - 1) + S11 * (q[i - 1] - q[i]) + q[i]
                    xbr = _edge01(q, dx, i, 0, minmax)
                elif kind == 2:
                    xbl = _edge01(q, dx, i, 1, minmax)
                    xbr = S15 * q[i] + S11 * q[i + 1] - S14 * dm_at(i + 1)
                elif kind == 3:
                    xbl = S15 * q[i - 1] + S11 * q[i] - S14 * dm_full[i]
                    xbr = al_full[i + 1]
                elif kind == 4:
                    xbl = al_full[i]
                    xbr = S15 * q[i + 1] + S11 * q[i] + S14 * dm_full[i]
                elif kind == 5:
                    xbl = S15 * q[i] + S11 * q[i - 1] + S14 * dm_at(i - 1)
                    xbr = _edge01(q, dx, i, 0, minmax)
                else:
                    xbl = _edge01(q, dx, i, 1, minmax)
                    xbr = S11 * (q[i + 1] - q[i]) - S14 * dm_at(i + 1) + q[i]
                a_l = xbl - q[i]
                a_r = xbr - q[i]
                # pert_ppm_standard_constraint_fcn (ppm.py:22-36)
                da1 = a_l - a_r
                da2 = da1 ** 2
                a6da = 3.0 * (a_l + a_r) * da1
                neg = a_l * a_r < 0.0
                new_r = np.where(neg & (a6da < -da2), -2.0 * a_l, a_r)
                new_l = np.where(neg & ~(a6da < -da2) & (a6da > da2), -2.0 * a_r, a_l)
                p = i - cells[0]
                bl[p] = np.where(neg, new_l, 0.0)
                br[p] = np.where(neg, new_r, 0.0)
        b0 = bl + br
        mask = 1.0
    ii = np.arange(i0, i1 + 1)
    cc = c[ii]
    fx1 = np.where(cc > 0.0, (1.0 - cc) * (br[:-1] - cc * b0[:-1]), (1.0 + cc) * (bl[1:] + cc * b0[1:]))   # fx1_fn
    return np.where(cc > 0.0, q[ii - 1] + fx1 * mask, q[ii] + fx1 * mask)                                   # apply_flux


def xppm(ix, q, c, dxa, ord_, i0, i1, j0, j1, K, out):
    out[i0 : i1 + 1, j0 : j1 + 1, K] = ppm_flux(q[:, j0 : j1 + 1, K], c[:, j0 : j1 + 1, K], dxa[:, j0 : j1 + 1], ord_,
                                                ix.west, ix.east, ix.isc, ix.iec, i0, i1)


def yppm(ix, q, c, dya, ord_, i0, i1, j0, j1, K, out):
    r = ppm_flux(q[i0 : i1 + 1, :, K].transpose(1, 0, 2), c[i0 : i1 + 1, :, K].transpose(1, 0, 2), dya[i0 : i1 + 1, :].T,
                 ord_, ix.south, ix.north, ix.jsc, ix.jec, j0, j1)
    out[i0 : i1 + 1, j0 : j1 + 1, K] = r.transpose(1, 0, 2)


# ---------------------------------------------------------------------------------------------
def calc_damp(damp_c, da_min, nord):
    return (np.asarray(damp_c) * da_min) ** (np.asarray(nord) + 1)


def delnflux_nosg(ix: Idx, g, q, fx2, fy2, damp, nord, nk, d2=None, mass=None):
    """DelnFluxNoSG.__call__ (delnflux.py:1209-1261).  damp, nord: arrays [>= nk].  fx2, fy2 (and d2) in place."""
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    nmax = int(max(nord[:nk]))
    if d2 is None:
        d2 = np.zeros_like(q)
    del6_u, del6_v, rarea = g["damp_del6_u"], g["damp_del6_v"], g["rarea"]
    for k in range(nk):
        K = slice(k, k + 1)
        hi = nord[k] > 0
        r = nmax if hi else 0
        # d2_damp_interval / copy_stencil_interval (:165-249)
        if hi:
            si, sj = sl(isc - 1 - nmax, iec + 1 + nmax), sl(jsc - 1 - nmax, jec + 1 + nmax)
        else:
            si, sj = sl(isc - 1, iec + 1), sl(jsc - 1, jec + 1)
        d2[si, sj, K] = q[si, sj, K] if mass is not None else damp[k] * q[si, sj, K]
        if hi:
            copy_corners_x(ix, d2, K)
        si, sj = sl(isc - r, iec + 1 + r), sl(jsc - r, jec + r)
        fx2[si, sj, K] = del6_v[si, sj, None] * (d2[si.start - 1 : si.stop - 1, sj, K] - d2[si, sj, K])
        if hi:
            copy_corners_y(ix, d2, K)
        si, sj = sl(isc - r, iec + r), sl(jsc - r, jec + 1 + r)
        fy2[si, sj, K] = del6_u[si, sj, None] * (d2[si, sj.start - 1 : sj.stop - 1, K] - d2[si, sj, K])
        if not hi:
            continue
        for n in range(nmax):
            nt = nmax - 1 - n
            si, sj = sl(isc - nt - 1, iec + nt + 1), sl(jsc - nt - 1, jec + nt + 1)
            d2[si, sj, K] = (fx2[si, sj, K] - fx2[si.start + 1 : si.stop + 1, sj, K] + fy2[si, sj, K]
                             - fy2[si, sj.start + 1 : sj.stop + 1, K]) * rarea[si, sj, None]
            copy_corners_x(ix, d2, K)
            si, sj = sl(isc - nt, iec + nt + 1), sl(jsc - nt, jec + nt)
            fx2[si, sj, K] = -del6_v[si, sj, None] * (d2[si.start - 1 : si.stop - 1, sj, K] - d2[si, sj, K])
            copy_corners_y(ix, d2, K)
            si, sj = sl(isc - nt, iec + nt), sl(jsc - nt, jec + nt + 1)
            fy2[si, sj, K] = -del6_u[si, sj, None] * (d2[si, sj.start - 1 : sj.stop - 1, K] - d2[si, sj, K])


def fvtp2d(ix: Idx, g, q, crx, cry, xfx, yfx, fx, fy, hord, nk, x_mass_flux=None, y_mass_flux=None, mass=None,
           nord=None, damp_c=None, da_min=None):
    """FiniteVolumeTransport.__call__ (fvtp2d.py:235-346).  q's cube-corner halos are rewritten in place, as in
    the reference; fx, fy in place."""
    K = slice(0, nk)
    isc, iec, jsc, jec, ied, jed = ix.isc, ix.iec, ix.jsc, ix.jec, ix.ied, ix.jed
    area = g["area"]
    ord_outer, ord_inner = hord, (8 if hord == 10 else hord)
    xu = xfx if x_mass_flux is None else x_mass_flux
    yu = yfx if y_mass_flux is None else y_mass_flux
    fy_in = np.zeros_like(q)
    fx_in = np.zeros_like(q)
    q_i = np.zeros_like(q)
    q_j = np.zeros_like(q)
    KC = slice(0, ix.nz + 1)  # CopyCorners runs on nz+1 levels (corners.py:25-27)
    copy_corners_y(ix, q, KC)
    yppm(ix, q, cry, g["dya"], ord_inner, 0, ied, jsc, jec + 1, K, fy_in)
    si, sj = sl(0, ied), sl(jsc, jec)
    fyy = yfx[:, :, K] * fy_in[:, :, K]
    q_i[si, sj, K] = (q[si, sj, K] * area[si, sj, None] + fyy[si, sj] - fyy[si, sj.start + 1 : sj.stop + 1]) / (
        area[si, sj, None] + yfx[si, sj, K] - yfx[si, sj.start + 1 : sj.stop + 1, K])
    outer_x = np.zeros_like(q)
    xppm(ix, q_i, crx, g["dxa"], ord_outer, isc, iec + 1, jsc, jec, K, outer_x)
    copy_corners_x(ix, q, KC)
    xppm(ix, q, crx, g["dxa"], ord_inner, isc, iec + 1, 0, jed, K, fx_in)
    si, sj = sl(isc, iec), sl(0, jed)
    fx1 = xfx[:, :, K] * fx_in[:, :, K]
    q_j[si, sj, K] = (q[si, sj, K] * area[si, sj, None] + fx1[si, sj] - fx1[si.start + 1 : si.stop + 1, sj]) / (
        area[si, sj, None] + xfx[si, sj, K] - xfx[si.start + 1 : si.stop + 1, sj, K])
    outer_y = np.zeros_like(q)
    yppm(ix, q_j, cry, g["dya"], ord_outer, isc, iec, jsc, jec + 1, K, outer_y)
    si, sj = sl(isc, iec + 1), sl(jsc, jec)
    fx[si, sj, K] = 0.5 * (outer_x[si, sj, K] + fx_in[si, sj, K]) * xu[si, sj, K]
    si, sj = sl(isc, iec), sl(jsc, jec + 1)
    fy[si, sj, K] = 0.5 * (outer_y[si, sj, K] + fy_in[si, sj, K]) * yu[si, sj, K]
    if nord is not None and damp_c is not None and not (np.asarray(damp_c) <= 1e-4).all():
        damp = calc_damp(damp_c, da_min, nord)
        fx2 = np.zeros_like(q)
        fy2 = np.zeros_like(q)
        delnflux_nosg(ix, g, q, fx2, fy2, damp, nord, nk, mass=mass)
        si, sj = sl(isc, iec + 1), sl(jsc, jec + 1)
        if mass is None:
            fx[si, sj, K] = fx[si, sj, K] + fx2[si, sj, K]
            fy[si, sj, K] = fy[si, sj, K] + fy2[si, sj, K]
        else:
            dk = damp[None, None, :nk]
            fx[si, sj, K] = fx[si, sj, K] + 0.5 * dk * (mass[si.start - 1 : si.stop - 1, sj, K] + mass[si, sj, K]) * fx2[si, sj, K]
            fy[si, sj, K] = fy[si, sj, K] + 0.5 * dk * (mass[si, sj.start - 1 : sj.stop - 1, K] + mass[si, sj, K]) * fy2[si, sj, K]
