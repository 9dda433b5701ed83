"""Oracle: DGridShallowWaterLagrangianDynamics.__call__ (fv3core/pace/fv3core/stencils/d_sw.py:935-1237) — test
infrastructure.  Composition of the restated pieces (fxadv, fv_tp_2d, delnflux, divergence damping) with the
remaining stencils written point by point (vectorised over levels): flux_capacitor (:29-50), heat_diss (:53-103),
apply_fluxes / apply_pt_delp_fluxes / adjust_w_and_qcon (:106-160, 331-346), compute_kinetic_energy (:204-298) with
advect_u_along_x / advect_v_along_y (xtp_u.py:9-91, ytp_v.py), compute_vorticity (:301-328), u_and_v_from_ke
(:439-477), vort_differencing + heat_source_from_vorticity_damping (:349-577), update_u_and_v (:582-608).
hord_mt < 8 only."""
import numpy as np

from .divergence_damping import divergence_damping
from .fvtp2d import calc_damp, delnflux_nosg, fvtp2d
from .fxadv import fv_prep
from .indexing import Idx

P1, P2 = 7.0 / 12.0, -1.0 / 12.0
C1, C2, C3 = -2.0 / 14.0, 11.0 / 14.0, 5.0 / 14.0


def _al(q, dx, i, lo, hi, start, end):
    """compute_al for ord < 8 (xppm.py:148-181); q(n), dx(n) accessors along the sweep."""
    if (lo and i == start - 1) or (hi and i == end):
        return C1 * q(i - 2) + C2 * q(i - 1) + C3 * q(i)
    if (lo and i == start) or (hi and i == end + 1):
        return 0.5 * (((2.0 * dx(i - 1) + dx(i - 2)) * q(i - 1) - dx(i - 1) * q(i - 2)) / (dx(i - 2) + dx(i - 1))
                      + ((2.0 * dx(i) + dx(i + 1)) * q(i) - dx(i) * q(i + 1)) / (dx(i) + dx(i + 1)))
    if (lo and i == start + 1) or (hi and i == end + 2):
        return C3 * q(i - 1) + C2 * q(i) + C1 * q(i + 1)
    return P1 * (q(i - 1) + q(i)) + P2 * (q(i - 2) + q(i + 1))


def _advect_along(mord, q, dx, zero, ub, cfl, i, lo, hi, start, end):
    al0, al1, al2 = (_al(q, dx, n, lo, hi, start, end) for n in (i - 1, i, i + 1))
    ql, qr = q(i - 1), q(i)
    bl_l, br_l, bl_r, br_r = al0 - ql, al1 - ql, al1 - qr, al2 - qr
    if zero(i - 1):
        bl_l = br_l = np.zeros_like(ql)
    if zero(i):
        bl_r = br_r = np.zeros_like(qr)
    b0_l, b0_r = bl_l + br_l, bl_r + br_r
    fx0 = np.where(cfl > 0.0, (1.0 - cfl) * (br_l - cfl * b0_l), (1.0 + cfl) * (bl_r + cfl * b0_r))
    if mord == 5:
        s_l, s_r = bl_l * br_l < 0, bl_r * br_r < 0
    else:
        s_l, s_r = (3.0 * np.abs(b0_l)) < np.abs(bl_l - br_l), (3.0 * np.abs(b0_r)) < np.abs(bl_r - br_r)
    mask = np.where(s_l | s_r, 1.0, 0.0)
    return np.where(ub > 0.0, ql + fx0 * mask, qr + fx0 * mask)


def d_sw(ix: Idx, g, a, col, cfg):
    """`a`: argument name -> [i, j, k] array of the reference call (updated in place); col: column namelist
    (get_column_namelist, d_sw.py:611-683); cfg: the d_grid_shallow_water namelist values."""
    nz = ix.nz
    isc, iec, jsc, jec, ied, jed = ix.isc, ix.iec, ix.jsc, ix.jec, ix.ied, ix.jed
    W, E, S, N = ix.west, ix.east, ix.south, ix.north
    K = slice(0, nz)
    dt = float(a["dt"])
    delp, pt, u, v, w, uc, vc = a["delp"], a["pt"], a["u"], a["v"], a["w"], a["uc"], a["vc"]
    ua, va, q_con = a["ua"], a["va"], a["q_con"]
    crx, cry, xfx, yfx = a["crx"], a["cry"], a["xfx"], a["yfx"]
    da_min, da_min_c = float(g["damp_da_min"]), float(g["damp_da_min_c"])
    z = lambda: np.zeros_like(delp)  # noqa: E731
    ucc, vcc = z(), z()
    fv_prep(ix, g, uc, vc, crx, cry, xfx, yfx, ucc, vcc, dt)
    fx, fy = z(), z()
    fvtp2d(ix, g, delp, crx, cry, xfx, yfx, fx, fy, cfg.hord_dp, nz, nord=col["nord_v"], damp_c=col["damp_vt"], da_min=da_min)
    F = (slice(0, ied + 1), slice(0, jed + 1), K)
    a["cx"][F] = a["cx"][F] + crx[F]
    a["cy"][F] = a["cy"][F] + cry[F]
    a["mfx"][F] = a["mfx"][F] + fx[F]
    a["mfy"][F] = a["mfy"][F] + fy[F]
    fx2, fy2 = z(), z()
    delnflux_nosg(ix, g, w, fx2, fy2, calc_damp(col["damp_w"], da_min_c, col["nord_w"]), col["nord_w"], nz)
    ci, cj = slice(isc, iec + 1), slice(jsc, jec + 1)
    ci1, cj1 = slice(isc + 1, iec + 2), slice(jsc + 1, jec + 2)
    ra = g["rarea"][ci, cj, None]
    damped_w = (np.asarray(col["damp_w"])[:nz] > 1e-5)[None, None, :]
    dw = (fx2[ci, cj, K] - fx2[ci1, cj, K] + fy2[ci, cj, K] - fy2[ci, cj1, K]) * ra
    dd8 = np.asarray(col["ke_bg"])[None, None, :nz] * abs(dt)
    heat_s = np.where(damped_w, dd8 - dw * (w[ci, cj, K] + 0.5 * dw), 0.0)
    a["diss_est"][ci, cj, K] = heat_s
    gxw, gyw, gxq, gyq, gxp, gyp = z(), z(), z(), z(), z(), z()
    fvtp2d(ix, g, w, crx, cry, xfx, yfx, gxw, gyw, cfg.hord_vt, nz, x_mass_flux=fx, y_mass_flux=fy)
    fvtp2d(ix, g, q_con, crx, cry, xfx, yfx, gxq, gyq, cfg.hord_dp, nz, x_mass_flux=fx, y_mass_flux=fy, mass=delp,
           nord=col["nord_t"], damp_c=col["damp_t"], da_min=da_min)
    fvtp2d(ix, g, pt, crx, cry, xfx, yfx, gxp, gyp, cfg.hord_tm, nz, x_mass_flux=fx, y_mass_flux=fy, mass=delp,
           nord=col["nord_v"], damp_c=col["damp_vt"], da_min=da_min)
    div = lambda gx, gy: (gx[ci, cj, K] - gx[ci1, cj, K] + gy[ci, cj, K] - gy[ci, cj1, K]) * ra  # noqa: E731
    dp0 = delp[ci, cj, K].copy()
    wv = w[ci, cj, K] * dp0 + div(gxw, gyw)
    qc = q_con[ci, cj, K] * dp0 + div(gxq, gyq)
    ptv = pt[ci, cj, K] * dp0 + div(gxp, gyp)
    dp1 = dp0 + div(fx, fy)
    ptv = ptv / dp1
    wv = wv / dp1
    wv = np.where(damped_w, wv + dw, wv)
    qc = qc / dp1
    delp[ci, cj, K], pt[ci, cj, K], w[ci, cj, K], q_con[ci, cj, K] = dp1, ptv, wv, qc
    # kinetic energy on cell corners
    mord = abs(cfg.hord_mt)
    ke = z()
    dx, dy, rdx, rdy = g["dx"], g["dy"], g["rdx"], g["rdy"]
    for i in range(isc, iec + 2):
        for j in range(jsc, jec + 2):
            ie_ = (W and i == isc) or (E and i == iec + 1)
            je_ = (S and j == jsc) or (N and j == jec + 1)
            ub_cov = 0.5 * (uc[i, j - 1, K] + uc[i, j, K])
            vb_cov = 0.5 * (vc[i - 1, j, K] + vc[i, j, K])
            ub = (ub_cov - vb_cov * g["cosa"][i, j]) * g["rsina"][i, j]
            vb = (vb_cov - ub_cov * g["cosa"][i, j]) * g["rsina"][i, j]
            if je_:
                ub = 0.25 * (-ucc[i, j - 2, K] + 3.0 * (ucc[i, j - 1, K] + ucc[i, j, K]) - ucc[i, j + 1, K])
            if ie_:
                ub = 0.5 * (ucc[i, j - 1, K] + ucc[i, j, K])
            if ie_:
                vb = 0.25 * (-vcc[i - 2, j, K] + 3.0 * (vcc[i - 1, j, K] + vcc[i, j, K]) - vcc[i + 1, j, K])
            if je_:
                vb = 0.5 * (vcc[i - 1, j, K] + vcc[i, j, K])
            if ie_ and je_:
                if i == isc and j == jsc:
                    io1, jo1, io2, vsign = 0, 0, -1, 1
                elif i != isc and j == jsc:
                    io1, jo1, io2, vsign = -1, 0, 0, -1
                elif i != isc and j != jsc:
                    io1, jo1, io2, vsign = -1, -1, 0, 1
                else:
                    io1, jo1, io2, vsign = 0, -1, -1, -1
                dt6 = dt / 6.0
                u0, um, v0, vm = u[i, j, K], u[i - 1, j, K], v[i, j, K], v[i, j - 1, K]
                ut0, utm, vt0, vtm = ucc[i, j, K], ucc[i, j - 1, K], vcc[i, j, K], vcc[i - 1, j, K]
                kev = dt6 * ((ut0 + utm) * ((io1 + 1) * u0 - (io1 * um)) + (vt0 + vtm) * ((jo1 + 1) * v0 - (jo1 * vm))
                             + (((jo1 + 1) * ut0 - (jo1 * utm)) + vsign * ((io1 + 1) * vt0 - (io1 * vtm))) * ((io2 + 1) * u0 - (io2 * um)))
            else:
                qu = lambda n: u[n, j, K]  # noqa: E731
                dxe = lambda n: dx[n, j]  # noqa: E731
                zx = lambda n: je_ and ((W and (n == isc - 1 or n == isc)) or (E and (n == iec or n == iec + 1)))  # noqa: E731
                cflx = np.where(ub > 0, ub * dt * rdx[i - 1, j], ub * dt * rdx[i, j])
                adv_u = _advect_along(mord, qu, dxe, zx, ub, cflx, i, W, E, isc, iec)
                qv = lambda n: v[i, n, K]  # noqa: E731
                dye = lambda n: dy[i, n]  # noqa: E731
                zy = lambda n: ie_ and ((S and (n == jsc - 1 or n == jsc)) or (N and (n == jec or n == jec + 1)))  # noqa: E731
                cfly = np.where(vb > 0, vb * dt * rdy[i, j - 1], vb * dt * rdy[i, j])
                adv_v = _advect_along(mord, qv, dye, zy, vb, cfly, j, S, N, jsc, jec)
                kev = 0.5 * dt * (ub * adv_u + vb * adv_v)
            ke[i, j, K] = kev
    # relative vorticity on the A grid, full domain
    fi, fj = slice(0, ied + 1), slice(0, jed + 1)
    fi1, fj1 = slice(1, ied + 2), slice(1, jed + 2)
    rdy_tmp = (g["rarea"][fi, fj] * dx[fi, fj])[:, :, None]
    rdx_tmp = (g["rarea"][fi, fj] * dy[fi, fj])[:, :, None]
    vort_a = z()
    vort_a[fi, fj, K] = (u[fi, fj, K] - u[fi, fj1, K] * dx[fi, fj1, None] / dx[fi, fj, None]) * rdy_tmp + (
        v[fi1, fj, K] * dy[fi1, fj, None] / dy[fi, fj, None] - v[fi, fj, K]) * rdx_tmp
    vort_b = z()
    nord_col = np.asarray(col["nord"])
    k0 = int(np.argmax(nord_col > 0)) if (nord_col > 0).any() else nz
    divergence_damping(ix, g, u, v, va, vort_b, ua, a["divgd"], vc, uc, a["delpc"], ke, vort_a, dt, np.asarray(col["d2_divg"]),
                       k0, int(nord_col.max()), cfg.dddmp, cfg.d4_bg)
    abs_vort = z()
    abs_vort[fi, fj, K] = vort_a[fi, fj, K] + g["fC_agrid"][fi, fj, None]
    fx, fy = z(), z()
    fvtp2d(ix, g, abs_vort, crx, cry, xfx, yfx, fx, fy, cfg.hord_vt, nz)
    bi, bj = slice(isc, iec + 2), slice(jsc, jec + 2)
    ui, uj = slice(isc, iec + 1), slice(jsc, jec + 2)
    u[ui, uj, K] = u[ui, uj, K] * dx[ui, uj, None] + ke[ui, uj, K] - ke[isc + 1 : iec + 2, uj, K] + fy[ui, uj, K]
    vi, vj = slice(isc, iec + 2), slice(jsc, jec + 1)
    v[vi, vj, K] = v[vi, vj, K] * dy[vi, vj, None] + ke[vi, vj, K] - ke[vi, jsc + 1 : jec + 2, K] - fx[vi, vj, K]
    ut, vt = z(), z()
    delnflux_nosg(ix, g, vort_a, ut, vt, calc_damp(col["damp_vt"], da_min_c, col["nord_v"]), col["nord_v"], nz)
    d_con = np.asarray(col["d_con"])[:nz]
    dc = (d_con > 1e-5)[None, None, :]
    ubt = lambda si, sj: (np.where(dc, vort_b[si, sj, K] - vort_b[si.start + 1 : si.stop + 1, sj, K], 0.0) + vt[si, sj, K]) * rdx[si, sj, None]  # noqa: E731
    vbt = lambda si, sj: (np.where(dc, vort_b[si, sj, K] - vort_b[si, sj.start + 1 : sj.stop + 1, K], 0.0) - ut[si, sj, K]) * rdy[si, sj, None]  # noqa: E731
    ub0, ub1 = ubt(ci, cj), ubt(ci, cj1)
    vb0, vb1 = vbt(ci, cj), vbt(ci1, cj)
    fy0, fy1 = u[ci, cj, K] * rdx[ci, cj, None], u[ci, cj1, K] * rdx[ci, cj1, None]
    fx0, fx1 = v[ci, cj, K] * rdy[ci, cj, None], v[ci1, cj, K] * rdy[ci1, cj, None]
    gy0, gy1, gx0, gx1 = fy0 * ub0, fy1 * ub1, fx0 * vb0, fx1 * vb1
    u2, du2, v2, dv2 = fy0 + fy1, ub0 + ub1, fx0 + fx1, vb0 + vb1
    dampterm = g["rsin2"][ci, cj, None] * 0.25 * ((ub0 * ub0 + ub1 * ub1 + vb0 * vb0 + vb1 * vb1) + 2.0 * (gy0 + gy1 + gx0 + gx1)
                                                 - g["cosa_s"][ci, cj, None] * (u2 * dv2 + v2 * du2 + du2 * dv2))
    hs = np.where(dc, delp[ci, cj, K] * (heat_s - d_con[None, None, :] * dampterm), heat_s)
    if cfg.d_con > 1e-5:
        a["heat_source"][ci, cj, K] = a["heat_source"][ci, cj, K] + hs
    dv = (np.asarray(col["damp_vt"])[:nz] > 1e-5)[None, None, :]
    u[ui, uj, K] = np.where(dv, u[ui, uj, K] + vt[ui, uj, K], u[ui, uj, K])
    v[vi, vj, K] = np.where(dv, v[vi, vj, K] - ut[vi, vj, K], v[vi, vj, K])
    del bi, bj
