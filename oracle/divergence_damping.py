"""Oracle: DivergenceDamping.__call__ (fv3core/pace/fv3core/stencils/divergence_damping.py:482-632) — test
infrastructure.  Point-by-point Python (vectorised over levels): second-order damping on the levels without
hyperdiffusion (:21-118, 504-548), `nord` Laplacian iterations of the divergence with the cube-corner fills of the
B-grid field (fill_corners_bgrid_x / _y, stencils/corners.py:591-702) and of the D-grid vector
(fill_corners_dgrid_defn, :987-1151) evaluated where they are read (:566-589), the Smagorinsky term on the B-grid
vorticity (a2b_ord4) and the high-order damping (:590-632)."""
import numpy as np

from .a2b import a2b_ord4
from .indexing import Idx


def divergence_damping(ix: Idx, g, u, v, va, vort_b, ua, divg_d, vc, uc, delpc, ke, vort_a, dt, d2_bg, k0, nord,
                       dddmp, d4_bg):
    nz = ix.nz
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    W, E, S, N = ix.west, ix.east, ix.south, ix.north
    da_min_c = float(g["damp_da_min_c"])
    ic, jc = iec + 1, jec + 1
    if k0 > 0:
        K = slice(0, k0)

        def ucd(i, j):
            if (S and j == jsc) or (N and j == jec + 1):
                uco = np.where(vc[i, j, K] > 0, u[i, j, K] * g["sin_sg4"][i, j - 1], u[i, j, K] * g["sin_sg2"][i, j])
            else:
                uco = (u[i, j, K] - 0.5 * (va[i, j - 1, K] + va[i, j, K]) * g["cosa_v"][i, j]) * g["sina_v"][i, j]
            return uco * g["dyc"][i, j]

        def vcd(i, j):
            if (W and i == isc) or (E and i == iec + 1):
                vco = np.where(uc[i, j, K] > 0, v[i, j, K] * g["sin_sg3"][i - 1, j], v[i, j, K] * g["sin_sg1"][i, j])
            else:
                vco = (v[i, j, K] - 0.5 * (ua[i - 1, j, K] + ua[i, j, K]) * g["cosa_u"][i, j]) * g["sina_u"][i, j]
            return vco * g["dxc"][i, j]

        for i in range(isc, iec + 2):
            for j in range(jsc, jec + 2):
                vm, v0, um, u0 = vcd(i, j - 1), vcd(i, j), ucd(i - 1, j), ucd(i, j)
                d = vm - v0 + um - u0
                ci = (W and i == isc) or (E and i == iec + 1)
                if ci and S and j == jsc:
                    d = d - vm
                if ci and N and j == jec + 1:
                    d = d + v0
                d = g["rarea_c"][i, j] * d
                delpc[i, j, K] = d
                damp = da_min_c * np.maximum(d2_bg[K], np.minimum(0.2, dddmp * np.abs(d * dt)))
                vo = damp * d
                vort_b[i, j, K] = vo
                ke[i, j, K] = ke[i, j, K] + vo
    K = slice(k0, nz)
    ci_, cj_ = slice(isc, iec + 2), slice(jsc, jec + 2)
    delpc[ci_, cj_, K] = divg_d[ci_, cj_, K]

    def corner(i, j, lo_i, hi_i, lo_j, hi_j):
        xw, xe, ys, yn = i < lo_i, i > hi_i, j < lo_j, j > hi_j
        return (xw or xe) and (ys or yn) and (W if xw else E) and (S if ys else N)

    dold = divg_d.copy()
    for n in range(nord):
        nt = nord - (n + 1)
        fillc = n + 1 != nord

        def dgx(i, j):
            if fillc and corner(i, j, isc, ic, jsc, jc):
                a = isc - i if i < isc else i - ic
                b = jsc - j if j < jsc else j - jc
                i, j = (isc - b if i < isc else ic + b), (jsc + a if j < jsc else jc - a)
            return dold[i, j, K]

        def dgy(i, j):
            if fillc and corner(i, j, isc, ic, jsc, jc):
                a = isc - i if i < isc else i - ic
                b = jsc - j if j < jsc else j - jc
                i, j = (isc + b if i < isc else ic - b), (jsc - a if j < jsc else jc + a)
            return dold[i, j, K]

        vc_raw = lambda i, j: (dgx(i + 1, j) - dgx(i, j)) * g["damp_divg_u"][i, j]  # noqa: E731
        uc_raw = lambda i, j: (dgy(i, j + 1) - dgy(i, j)) * g["damp_divg_v"][i, j]  # noqa: E731

        def vc_at(i, j):
            if fillc and corner(i, j, isc, iec, jsc, jec + 1):
                xw, ys = i < isc, j < jsc
                a = isc - i if xw else i - iec
                b = jsc - j if ys else j - (jec + 1)
                sg = -1.0 if xw == ys else 1.0
                return sg * uc_raw(isc - b if xw else iec + 1 + b, jsc + a - 1 if ys else jec + 1 - a)
            return vc_raw(i, j)

        def uc_at(i, j):
            if fillc and corner(i, j, isc, iec + 1, jsc, jec):
                xw, ys = i < isc, j < jsc
                a = isc - i if xw else i - (iec + 1)
                b = jsc - j if ys else j - jec
                sg = -1.0 if xw == ys else 1.0
                return sg * vc_raw(isc + b - 1 if xw else iec + 1 - b, jsc - a if ys else jec + 1 + a)
            return uc_raw(i, j)

        dnew = dold.copy()
        for i in range(isc - nt, iec + nt + 2):
            for j in range(jsc - nt, jec + nt + 2):
                ucm, uc0, vcm, vc0 = uc_at(i, j - 1), uc_at(i, j), vc_at(i - 1, j), vc_at(i, j)
                d = ucm - uc0 + vcm - vc0
                ci = (W and i == isc) or (E and i == iec + 1)
                if ci and S and j == jsc:
                    d = d - ucm
                if ci and N and j == jec + 1:
                    d = d + uc0
                dnew[i, j, K] = d * g["rarea_c"][i, j]
        dold = dnew
    divg_d[ci_, cj_, K] = dold[ci_, cj_, K]
    absdt = abs(dt)
    dd8 = (da_min_c * d4_bg) ** (nord + 1)
    if dddmp < 1e-5:
        vb = np.zeros_like(divg_d)
    else:
        vb = a2b_ord4(ix, g, vort_a, k0, nz)
    for i in range(isc, iec + 2):
        for j in range(jsc, jec + 2):
            if dddmp < 1e-5:
                vo = np.zeros(nz - k0)
            else:
                dp = delpc[i, j, K]
                vo = absdt * np.sqrt(dp * dp + vb[i, j, K] * vb[i, j, K])
            damp = da_min_c * np.maximum(d2_bg[K], np.minimum(0.2, dddmp * np.abs(vo)))
            vo = damp * delpc[i, j, K] + dd8 * divg_d[i, j, K]
            vort_b[i, j, K] = vo
            ke[i, j, K] = ke[i, j, K] + vo
