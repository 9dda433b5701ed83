"""Oracle: A-grid -> B-grid 4th-order interpolation and the non-hydrostatic pressure-gradient update (test
infrastructure).  Point-by-point Python (vectorised over levels) restatement of
AGrid2BGridFourthOrder.__call__ (fv3core/pace/fv3core/stencils/a2b_ord4.py:673-761): corner extrapolation
(:37-273), tile-edge values qout_x_edge / qout_y_edge (:286-311), ppm_volume_mean_x / _y (:416-450) with their
tile-edge forms (:314-413), a2b_interpolation (:453-481); and of NonHydrostaticPressureGradient.__call__
(nh_p_grad.py:190-255) with calc_u / calc_v (:23-112)."""
import numpy as np

from .indexing import Idx

C1, C2, B1, B2, A1, A2 = 2.0 / 3.0, -1.0 / 6.0, 7.0 / 12.0, -1.0 / 12.0, 9.0 / 16.0, -1.0 / 16.0


def _gcd(lon1, lat1, lon2, lat2):
    """great_circle_dist (a2b_ord4.py:37-41) on the unit sphere."""
    tb = np.sin((lat1 - lat2) / 2.0) ** 2.0
    ta = np.sin((lon1 - lon2) / 2.0) ** 2.0
    return np.arcsin(np.sqrt(tb + np.cos(lat1) * np.cos(lat2) * ta)) * 2.0


def a2b_ord4(ix: Idx, g, qin, k0, k1):
    """qout on the corner points [isc..iec+1] x [jsc..jec+1] for levels [k0, k1) (other points zero)."""
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    W, E, S, N = ix.west, ix.east, ix.south, ix.north
    dxa, dya = g["dxa"], g["dya"]
    K = slice(k0, k1)
    q = lambda i, j: qin[i, j, K]  # noqa: E731

    def qx(i, j):
        if W and i == isc:
            gi, go = dxa[i + 1, j] / dxa[i, j], dxa[i - 2, j] / dxa[i - 1, j]
            return 0.5 * (((2.0 + gi) * q(i, j) - q(i + 1, j)) / (1.0 + gi) + ((2.0 + go) * q(i - 1, j) - q(i - 2, j)) / (1.0 + go))
        if W and i == isc + 1:
            gi, go = dxa[i, j] / dxa[i - 1, j], dxa[i - 3, j] / dxa[i - 2, j]
            left = 0.5 * (((2.0 + gi) * q(i - 1, j) - q(i, j)) / (1.0 + gi) + ((2.0 + go) * q(i - 2, j) - q(i - 3, j)) / (1.0 + go))
            right = B2 * (q(i - 1, j) + q(i + 2, j)) + B1 * (q(i, j) + q(i + 1, j))
            return (3.0 * (gi * q(i - 1, j) + q(i, j)) - (gi * left + right)) / (2.0 + 2.0 * gi)
        if E and i == iec + 1:
            gi, go = dxa[i - 2, j] / dxa[i - 1, j], dxa[i + 1, j] / dxa[i, j]
            return 0.5 * (((2.0 + gi) * q(i - 1, j) - q(i - 2, j)) / (1.0 + gi) + ((2.0 + go) * q(i, j) - q(i + 1, j)) / (1.0 + go))
        if E and i == iec:
            gi, go = dxa[i - 1, j] / dxa[i, j], dxa[i + 2, j] / dxa[i + 1, j]
            right = 0.5 * (((2.0 + gi) * q(i, j) - q(i - 1, j)) / (1.0 + gi) + ((2.0 + go) * q(i + 1, j) - q(i + 2, j)) / (1.0 + go))
            left = B2 * (q(i - 3, j) + q(i, j)) + B1 * (q(i - 2, j) + q(i - 1, j))
            return (3.0 * (q(i - 1, j) + gi * q(i, j)) - (gi * right + left)) / (2.0 + 2.0 * gi)
        return B2 * (q(i - 2, j) + q(i + 1, j)) + B1 * (q(i - 1, j) + q(i, j))

    def qy(i, j):
        if S and j == jsc:
            gi, go = dya[i, j + 1] / dya[i, j], dya[i, j - 2] / dya[i, j - 1]
            return 0.5 * (((2.0 + gi) * q(i, j) - q(i, j + 1)) / (1.0 + gi) + ((2.0 + go) * q(i, j - 1) - q(i, j - 2)) / (1.0 + go))
        if S and j == jsc + 1:
            gi, go = dya[i, j] / dya[i, j - 1], dya[i, j - 3] / dya[i, j - 2]
            lower = 0.5 * (((2.0 + gi) * q(i, j - 1) - q(i, j)) / (1.0 + gi) + ((2.0 + go) * q(i, j - 2) - q(i, j - 3)) / (1.0 + go))
            upper = B2 * (q(i, j - 1) + q(i, j + 2)) + B1 * (q(i, j) + q(i, j + 1))
            return (3.0 * (gi * q(i, j - 1) + q(i, j)) - (gi * lower + upper)) / (2.0 + 2.0 * gi)
        if N and j == jec + 1:
            gi, go = dya[i, j - 2] / dya[i, j - 1], dya[i, j + 1] / dya[i, j]
            return 0.5 * (((2.0 + gi) * q(i, j - 1) - q(i, j - 2)) / (1.0 + gi) + ((2.0 + go) * q(i, j) - q(i, j + 1)) / (1.0 + go))
        if N and j == jec:
            gi, go = dya[i, j - 1] / dya[i, j], dya[i, j + 2] / dya[i, j + 1]
            lower = B2 * (q(i, j - 3) + q(i, j)) + B1 * (q(i, j - 2) + q(i, j - 1))
            upper = 0.5 * (((2.0 + gi) * q(i, j) - q(i, j - 1)) / (1.0 + gi) + ((2.0 + go) * q(i, j + 1) - q(i, j + 2)) / (1.0 + go))
            return (3.0 * (q(i, j - 1) + gi * q(i, j)) - (gi * upper + lower)) / (2.0 + 2.0 * gi)
        return B2 * (q(i, j - 2) + q(i, j + 1)) + B1 * (q(i, j - 1) + q(i, j))

    lon, lat, lona, lata = g["lon"], g["lat"], g["lon_agrid"], g["lat_agrid"]

    def edge_value(i, j):
        iw, ie = W and i == isc, E and i == iec + 1
        js, jn = S and j == jsc, N and j == jec + 1
        if (iw or ie) and (js or jn):
            di, dj = (1 if iw else -1), (1 if js else -1)
            i0, j0 = (i if iw else i - 1), (j if js else j - 1)
            arms = [((i0, j0), (i0 + di, j0 + dj)), ((i0 - di, j0), (i0 - 2 * di, j0 + dj)), ((i0, j0 - dj), (i0 + di, j0 - 2 * dj))]
            ext = []
            for p1, p2 in arms:                                    # extrap_corner (:43-56)
                x1 = _gcd(lona[p1], lata[p1], lon[i, j], lat[i, j])
                x2 = _gcd(lona[p2], lata[p2], lon[i, j], lat[i, j])
                ext.append(q(*p1) + x1 / (x2 - x1) * (q(*p1) - q(*p2)))
            e1, e2, e3 = ext
            total = (e1 + e3 + e2) if (js and ie) else (e1 + e2 + e3)
            return total * (1.0 / 3.0)
        if iw or ie:
            q2 = lambda jj: (q(i - 1, jj) * dxa[i, jj] + q(i, jj) * dxa[i - 1, jj]) / (dxa[i - 1, jj] + dxa[i, jj])  # noqa: E731
            ew = (g["edge_w"] if iw else g["edge_e"])
            ew = ew[j] if ew.ndim == 1 else ew[0, j]
            return ew * q2(j - 1) + (1.0 - ew) * q2(j)
        q1 = lambda ii: (q(ii, j - 1) * dya[ii, j] + q(ii, j) * dya[ii, j - 1]) / (dya[ii, j - 1] + dya[ii, j])  # noqa: E731
        es = (g["edge_s"] if js else g["edge_n"])
        es = es[i] if es.ndim == 1 else es[i, 0]
        return es * q1(i - 1) + (1.0 - es) * q1(i)

    out = np.zeros_like(qin)
    is_edge = lambda i, j: (W and i == isc) or (E and i == iec + 1) or (S and j == jsc) or (N and j == jec + 1)  # noqa: E731
    for i in range(isc, iec + 2):
        for j in range(jsc, jec + 2):
            if is_edge(i, j):
                out[i, j, K] = edge_value(i, j)
    for i in range(isc, iec + 2):
        for j in range(jsc, jec + 2):
            if is_edge(i, j):
                continue
            if S and j == jsc + 1:
                upper = A2 * (qx(i, j - 1) + qx(i, j + 2)) + A1 * (qx(i, j) + qx(i, j + 1))
                qxx = C1 * (qx(i, j - 1) + qx(i, j)) + C2 * (out[i, j - 1, K] + upper)
            elif N and j == jec:
                lower = A2 * (qx(i, j - 3) + qx(i, j)) + A1 * (qx(i, j - 2) + qx(i, j - 1))
                qxx = C1 * (qx(i, j - 1) + qx(i, j)) + C2 * (out[i, j + 1, K] + lower)
            else:
                qxx = A2 * (qx(i, j - 2) + qx(i, j + 1)) + A1 * (qx(i, j - 1) + qx(i, j))
            if W and i == isc + 1:
                right = A2 * (qy(i - 1, j) + qy(i + 2, j)) + A1 * (qy(i, j) + qy(i + 1, j))
                qyy = C1 * (qy(i - 1, j) + qy(i, j)) + C2 * (out[i - 1, j, K] + right)
            elif E and i == iec:
                left = A2 * (qy(i - 3, j) + qy(i, j)) + A1 * (qy(i - 2, j) + qy(i - 1, j))
                qyy = C1 * (qy(i - 1, j) + qy(i, j)) + C2 * (out[i + 1, j, K] + left)
            else:
                qyy = A2 * (qy(i - 2, j) + qy(i + 1, j)) + A1 * (qy(i - 1, j) + qy(i, j))
            out[i, j, K] = 0.5 * (qxx + qyy)
    return out


def nh_p_grad(ix: Idx, g, u, v, pp, gz, pk3, delp, dt, ptop, akap):
    """NonHydrostaticPressureGradient.__call__ (nh_p_grad.py:190-255): pp, pk3 (k >= 1), gz, delp to the B grid, top
    values (set_k0 :11-20), calc_u / calc_v (:23-112); u, v, pp, gz, pk3 in place."""
    nz = ix.nz
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    ci, cj = slice(isc, iec + 2), slice(jsc, jec + 2)
    ppb = a2b_ord4(ix, g, pp, 1, nz + 1)
    pk3b = a2b_ord4(ix, g, pk3, 1, nz + 1)
    gzb = a2b_ord4(ix, g, gz, 0, nz + 1)
    wk1 = a2b_ord4(ix, g, delp, 0, nz)
    ppb[ci, cj, 0] = 0.0
    pk3b[ci, cj, 0] = ptop ** akap
    pp[ci, cj, : nz + 1] = ppb[ci, cj, : nz + 1]
    pk3[ci, cj, : nz + 1] = pk3b[ci, cj, : nz + 1]
    gz[ci, cj, : nz + 1] = gzb[ci, cj, : nz + 1]
    rdx, rdy = g["rdx"], g["rdy"]
    K, K1 = slice(0, nz), slice(1, nz + 1)
    for i in range(isc, iec + 2):
        for j in range(jsc, jec + 2):
            if i <= iec:
                wk0 = pk3b[i, j, K1] - pk3b[i, j, K]
                wkx = pk3b[i + 1, j, K1] - pk3b[i + 1, j, K]
                a = gzb[i, j, K1] - gzb[i + 1, j, K]
                b = gzb[i, j, K] - gzb[i + 1, j, K1]
                du = dt / (wk0 + wkx) * (a * (pk3b[i + 1, j, K1] - pk3b[i, j, K]) + b * (pk3b[i, j, K1] - pk3b[i + 1, j, K]))
                u[i, j, K] = (u[i, j, K] + du + dt / (wk1[i, j, K] + wk1[i + 1, j, K]) * (
                    a * (ppb[i + 1, j, K1] - ppb[i, j, K]) + b * (ppb[i, j, K1] - ppb[i + 1, j, K]))) * rdx[i, j]
            if j <= jec:
                wk0 = pk3b[i, j, K1] - pk3b[i, j, K]
                wky = pk3b[i, j + 1, K1] - pk3b[i, j + 1, K]
                a = gzb[i, j, K1] - gzb[i, j + 1, K]
                b = gzb[i, j, K] - gzb[i, j + 1, K1]
                dv = dt / (wk0 + wky) * (a * (pk3b[i, j + 1, K1] - pk3b[i, j, K]) + b * (pk3b[i, j, K1] - pk3b[i, j + 1, K]))
                v[i, j, K] = (v[i, j, K] + dv + dt / (wk1[i, j, K] + wk1[i, j + 1, K]) * (
                    a * (ppb[i, j + 1, K1] - ppb[i, j, K]) + b * (ppb[i, j, K1] - ppb[i, j + 1, K]))) * rdy[i, j]
