"""Oracle: the small stencils of AcousticDynamics (fv3core/pace/fv3core/stencils/dyn_core.py:48-171) — test infra."""
import numpy as np

from .c_sw import _sh
from .constants import GRAV
from .indexing import Idx, sl


def gz_from_surface_height_and_thicknesses(ix: Idx, zs, delz, gz):
    """dyn_core.py:83-96, compute domain, nz+1 levels."""
    si, sj = sl(ix.isc, ix.iec), sl(ix.jsc, ix.jec)
    gz[si, sj, ix.nz] = zs[si, sj]
    for k in range(ix.nz - 1, -1, -1):
        gz[si, sj, k] = gz[si, sj, k + 1] - delz[si, sj, k]


def interface_pressure_from_toa_pressure_and_thickness(ix: Idx, delp, pem, ptop):
    """dyn_core.py:99-112, compute domain + 1 halo.  NOTE the reference adds delp at the SAME level k."""
    si, sj = sl(ix.isc - 1, ix.iec + 1), sl(ix.jsc - 1, ix.jec + 1)
    pem[si, sj, 0] = ptop
    for k in range(1, ix.nz):
        pem[si, sj, k] = pem[si, sj, k - 1] + delp[si, sj, k]


def p_grad_c(ix: Idx, rdxc, rdyc, uc, vc, delpc, pkc, gz, dt2):
    """p_grad_c_stencil (dyn_core.py:120-171), non-hydrostatic branch; uc, vc in place."""
    K = slice(0, ix.nz)
    K1 = slice(1, ix.nz + 1)
    si, sj = sl(ix.isc, ix.iec + 1), sl(ix.jsc, ix.jec + 1)
    wk = delpc
    gzm = _sh(gz, -1, 0, si, sj)
    pkm = _sh(pkc, -1, 0, si, sj)
    uc[si, sj, K] = uc[si, sj, K] + dt2 * rdxc[si, sj, None] / (_sh(wk, -1, 0, si, sj)[:, :, K] + wk[si, sj, K]) * (
        (gzm[:, :, K1] - gz[si, sj, K]) * (pkc[si, sj, K1] - pkm[:, :, K])
        + (gzm[:, :, K] - gz[si, sj, K1]) * (pkm[:, :, K1] - pkc[si, sj, K]))
    gzm = _sh(gz, 0, -1, si, sj)
    pkm = _sh(pkc, 0, -1, si, sj)
    vc[si, sj, K] = vc[si, sj, K] + dt2 * rdyc[si, sj, None] / (_sh(wk, 0, -1, si, sj)[:, :, K] + wk[si, sj, K]) * (
        (gzm[:, :, K1] - gz[si, sj, K]) * (pkc[si, sj, K1] - pkm[:, :, K])
        + (gzm[:, :, K] - gz[si, sj, K1]) * (pkm[:, :, K1] - pkc[si, sj, K]))


def compute_geopotential(ix: Idx, zh, gz):
    """dyn_core.py:115-117 on compute + 2 halo, nz+1 levels."""
    si, sj = sl(ix.isc - 2, ix.iec + 2), sl(ix.jsc - 2, ix.jec + 2)
    gz[si, sj, : ix.nz + 1] = zh[si, sj, : ix.nz + 1] * GRAV


def edge_pe(ix: Idx, pe, delp, ptop):
    """edge_pe (pe_halo.py:6-34): interface pressure in the 1-cell ring around the compute domain."""
    for i in range(ix.isc - 1, ix.iec + 2):
        for j in range(ix.jsc - 1, ix.jec + 2):
            if ix.isc <= i <= ix.iec and ix.jsc <= j <= ix.jec:
                continue
            pe[i, j, 0] = ptop
            for k in range(1, ix.nz + 1):
                pe[i, j, k] = pe[i, j, k - 1] + delp[i, j, k - 1]


def pk3_halo(ix: Idx, pk3, delp, ptop, akap):
    """PK3Halo.__call__ (pk3_halo.py:11-69): pk3 = pe ** akap in the 2-cell ring around the compute domain."""
    for i in range(ix.isc - 2, ix.iec + 3):
        for j in range(ix.jsc - 2, ix.jec + 3):
            if ix.isc <= i <= ix.iec and ix.jsc <= j <= ix.jec:
                continue
            p = ptop
            for k in range(1, ix.nz + 1):
                p = p + delp[i, j, k - 1]
                pk3[i, j, k] = p ** akap


def apply_diffusive_heating(ix: Idx, delp, delz, cappa, heat_source, pt, delt_time_factor):
    """apply_diffusive_heating (temperature_adjust.py:8-43), compute domain, pt in place."""
    from .constants import CV_AIR, RDG

    si, sj, K = sl(ix.isc, ix.iec), sl(ix.jsc, ix.jec), slice(0, ix.nz)
    pkz = (RDG * delp[si, sj, K] / delz[si, sj, K] * pt[si, sj, K]) ** (cappa[si, sj, K] / (1.0 - cappa[si, sj, K]))
    dtmp = heat_source[si, sj, K] / (CV_AIR * delp[si, sj, K])
    lim = np.full(ix.nz, delt_time_factor)
    lim[0] = delt_time_factor * 0.1
    lim[1] = delt_time_factor * 0.5
    mag = np.minimum(lim[None, None, :], np.abs(dtmp))
    deltmin = np.where(dtmp > 0, np.abs(mag), -np.abs(mag))
    pt[si, sj, K] = pt[si, sj, K] + deltmin / pkz


def del2cubed(ix: Idx, g, qdel, cd, nmax):
    """HyperdiffusionDamping.__call__ (del2cubed.py:165-194): corner_fill (:25-65), copy_corners_x / _y
    (stencils/corners.py:307-425), compute_zonal / meridional_flux (:15-22), update_q (:72-76); qdel in place."""
    from .fvtp2d import copy_corners_x, copy_corners_y

    K = slice(0, ix.nz)
    third = 1.0 / 3.0
    ntimes = int(min(3, nmax))
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    for n in range(ntimes):
        nt = ntimes - (n + 1)
        q = qdel.copy()                                                     # corner_fill
        qi = qdel
        for (i, j, pts) in (
                (isc, jsc, ((0, 0), (-1, 0), (0, -1))), (isc - 1, jsc, ((1, 0), (0, 0), (1, -1))),
                (isc, jsc - 1, ((0, 1), (-1, 1), (0, 0))),
                (iec, jsc, ((0, 0), (1, 0), (0, -1))), (iec + 1, jsc, ((-1, 0), (0, 0), (-1, -1))),
                (iec, jsc - 1, ((0, 1), (1, 1), (0, 0))),
                (iec, jec, ((0, 0), (1, 0), (0, 1))), (iec + 1, jec, ((-1, 0), (0, 0), (-1, 1))),
                (iec, jec + 1, ((0, -1), (1, -1), (0, 0))),
                (isc, jec, ((0, 0), (-1, 0), (0, 1))), (isc - 1, jec, ((1, 0), (0, 0), (1, 1))),
                (isc, jec + 1, ((0, -1), (-1, -1), (0, 0)))):
            (a, b, c_) = pts
            if not ((ix.west if i <= isc else ix.east) and (ix.south if j <= jsc else ix.north)):
                continue
            q[i, j, K] = (qi[i + a[0], j + a[1], K] + qi[i + b[0], j + b[1], K] + qi[i + c_[0], j + c_[1], K]) * third
        if nt > 0:
            copy_corners_x(ix, q, K)
        si, sj = sl(isc - nt, iec + nt + 1), sl(jsc - nt, jec + nt)
        fx = np.zeros_like(q)
        fx[si, sj, K] = g["damp_del6_v"][si, sj, None] * (_sh(q, -1, 0, si, sj)[:, :, K] - q[si, sj, K])
        if nt > 0:
            copy_corners_y(ix, q, K)
        si, sj = sl(isc - nt, iec + nt), sl(jsc - nt, jec + nt + 1)
        fy = np.zeros_like(q)
        fy[si, sj, K] = g["damp_del6_u"][si, sj, None] * (_sh(q, 0, -1, si, sj)[:, :, K] - q[si, sj, K])
        qdel[:, :, K] = q[:, :, K]
        si, sj = sl(isc - nt, iec + nt), sl(jsc - nt, jec + nt)
        qdel[si, sj, K] = qdel[si, sj, K] + cd * g["rarea"][si, sj, None] * (
            fx[si, sj, K] - _sh(fx, 1, 0, si, sj)[:, :, K] + fy[si, sj, K] - _sh(fy, 0, 1, si, sj)[:, :, K])


def ray_fast(ix: Idx, u, v, w, dp, pfull, dt, ptop, rf_cutoff, tau, hydrostatic=False):
    """RayleighDamping.__call__ (ray_fast.py:184-206) = ray_fast_wind_compute (:48-141): damping factors rf (:27-41),
    reference pressure p_ref summed over the nudged levels, momentum lost by u / v redistributed over those levels;
    u on (nx, ny+1), v on (nx+1, ny), w on (nx, ny); in place."""
    import math

    nz = ix.nz
    nudge = rf_cutoff + min(100.0, 10.0 * ptop)
    rf = np.ones(nz)
    damped = pfull[:nz] < rf_cutoff
    for k in range(nz):
        if damped[k]:
            val = dt / (tau * 86400.0) * math.sin(0.5 * math.pi * math.log(rf_cutoff / pfull[k]) / math.log(rf_cutoff / ptop)) ** 2
            rf[k] = 1.0 / (1.0 + val)
    p_ref = 0.0
    for k in range(nz):
        if pfull[k] < nudge:
            p_ref = dp[k] if k == 0 else p_ref + dp[k]
    for q, si, sj in ((u, sl(ix.isc, ix.iec), sl(ix.jsc, ix.jec + 1)), (v, sl(ix.isc, ix.iec + 1), sl(ix.jsc, ix.jec))):
        dm = None
        for k in range(nz):
            if damped[k]:
                d = (1.0 - rf[k]) * dp[k] * q[si, sj, k]
                dm = d if dm is None else dm + d
                q[si, sj, k] = q[si, sj, k] * rf[k]
        for k in range(nz):
            if pfull[k] < nudge and dm is not None:
                q[si, sj, k] = q[si, sj, k] + dm / p_ref
    if not hydrostatic:
        si, sj = sl(ix.isc, ix.iec), sl(ix.jsc, ix.jec)
        for k in range(nz):
            if damped[k]:
                w[si, sj, k] = w[si, sj, k] * rf[k]
