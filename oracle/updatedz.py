"""Oracle: geopotential height updates on the C grid (and, below, the D grid) — test infrastructure.

update_dz_c follows fv3core/pace/fv3core/stencils/updatedzc.py:15-207.
"""
import numpy as np

from .c_sw import _fill2_x, _fill2_y, _sh
from .constants import DZ_MIN
from .indexing import Idx, sl


def _interface_average(vel, dp0, nz):
    """p_weighted_average_{top,domain,bottom} (updatedzc.py:15-33): layer values -> nz+1 interfaces."""
    out = np.zeros(vel.shape[:2] + (nz + 1,))
    ratio = dp0[0] / (dp0[0] + dp0[1])
    out[:, :, 0] = vel[:, :, 0] + (vel[:, :, 0] - vel[:, :, 1]) * ratio
    for k in range(1, nz):
        int_ratio = 1.0 / (dp0[k - 1] + dp0[k])
        out[:, :, k] = (dp0[k] * vel[:, :, k - 1] + dp0[k - 1] * vel[:, :, k]) * int_ratio
    ratio = dp0[nz - 1] / (dp0[nz - 2] + dp0[nz - 1])
    out[:, :, nz] = vel[:, :, nz - 1] + (vel[:, :, nz - 1] - vel[:, :, nz - 2]) * ratio
    return out


def update_dz_c(ix: Idx, dp_ref, zs, area, ut, vt, gz, ws, dt):
    """UpdateGeopotentialHeightOnCGrid.__call__ (updatedzc.py:167-207); gz [.., nz+1] and ws updated in place."""
    nz = ix.nz
    KI = slice(0, nz + 1)
    isc, iec, jsc, jec = ix.isc, ix.iec, ix.jsc, ix.jec
    gz_x = gz.copy()
    gz_y = gz.copy()
    _fill2_x(ix, gz_x, KI)
    _fill2_y(ix, gz_y, KI)
    si, sj = sl(isc - 1, iec + 2), sl(jsc - 1, jec + 2)
    xfx = _interface_average(ut[si, sj, :nz], dp_ref, nz)
    yfx = _interface_average(vt[si, sj, :nz], dp_ref, nz)
    fx = xfx * np.where(xfx > 0.0, _sh(gz_x, -1, 0, si, sj)[:, :, KI], gz_x[si, sj, KI])
    fy = yfx * np.where(yfx > 0.0, _sh(gz_y, 0, -1, si, sj)[:, :, KI], gz_y[si, sj, KI])
    ci, cj = sl(isc - 1, iec + 1), sl(jsc - 1, jec + 1)
    ar = area[ci, cj, None]
    new = (gz[ci, cj, KI] * ar + fx[:-1, :-1] - fx[1:, :-1] + fy[:-1, :-1] - fy[:-1, 1:]) / (
        ar + xfx[:-1, :-1] - xfx[1:, :-1] + yfx[:-1, :-1] - yfx[:-1, 1:])
    rdt = 1.0 / dt
    ws[ci, cj] = (zs[ci, cj] - new[:, :, nz]) * rdt
    for k in range(nz - 1, -1, -1):
        kp1 = new[:, :, k + 1] + DZ_MIN
        new[:, :, k] = np.where(new[:, :, k] > kp1, new[:, :, k], kp1)
    gz[ci, cj, KI] = new


def _spline_constants(dp0, nz):
    """cubic_spline_interpolation_constants (updatedzd.py:129-154)."""
    gk = np.zeros(nz)
    beta = np.zeros(nz)
    gamma = np.zeros(nz)
    gk[0] = dp0[1] / dp0[0]
    beta[0] = gk[0] * (gk[0] + 0.5)
    gamma[0] = (1.0 + gk[0] * (gk[0] + 1.5)) / beta[0]
    gk[1:] = dp0[: nz - 1] / dp0[1:nz]
    for i in range(1, nz):
        beta[i] = 2.0 + 2.0 * gk[i] - gamma[i - 1]
        gamma[i] = gk[i] / beta[i]
    return gk, beta, gamma


def _to_interfaces(qc, gk, beta, gamma, nz):
    """cubic_spline_interpolation_from_layer_center_to_interfaces (updatedzd.py:157-196), whole horizontal domain."""
    qi = np.zeros(qc.shape[:2] + (nz + 1,))
    qi[:, :, 0] = (2.0 * gk[0] * (gk[0] + 1.0) * qc[:, :, 0] + qc[:, :, 1]) / beta[0]
    for k in range(1, nz):
        qi[:, :, k] = (3.0 * (qc[:, :, k - 1] + gk[k] * qc[:, :, k]) - qi[:, :, k - 1]) / beta[k]
    gl = gk[nz - 1]
    a_bot = 1.0 + gl * (gl + 1.5)
    xt1 = 2.0 * gl * (gl + 1.0)
    xt2 = gl * (gl + 0.5) - a_bot * gamma[nz - 1]
    qi[:, :, nz] = (xt1 * qc[:, :, nz - 1] + qc[:, :, nz - 2] - a_bot * qi[:, :, nz - 1]) / xt2
    for k in range(nz - 1, -1, -1):
        qi[:, :, k] = qi[:, :, k] - gamma[k] * qi[:, :, k + 1]
    return qi


def update_dz_d(ix: Idx, g, zs, zh, crx, cry, xfx, yfx, ws, dt, damp_vt, nord_v, hord_tm):
    """UpdateHeightOnDGrid.__call__ (updatedzd.py:283-356): spline interpolation of the Courant numbers / area fluxes to
    the interfaces, fv_tp_2d on the height (nz+1 levels), del-n damping fluxes with the raw damp_vt column,
    apply_height_fluxes (:70-126); zh and ws in place."""
    from .fvtp2d import delnflux_nosg, fvtp2d

    nz = ix.nz
    gk, beta, gamma = _spline_constants(g["dp_ref"], nz)
    crx_i, cry_i, xfx_i, yfx_i = (_to_interfaces(q, gk, beta, gamma, nz) for q in (crx, cry, xfx, yfx))
    fx = np.zeros_like(xfx_i)
    fy = np.zeros_like(xfx_i)
    fvtp2d(ix, g, zh, crx_i, cry_i, xfx_i, yfx_i, fx, fy, hord_tm, nz + 1)
    gx = np.zeros_like(xfx_i)
    gy = np.zeros_like(xfx_i)
    damp = np.append(np.asarray(damp_vt, dtype=np.float64)[:nz], 0.0)
    nord = np.append(np.asarray(nord_v, dtype=np.float64)[:nz], 0.0)
    delnflux_nosg(ix, g, zh, gx, gy, damp, nord, nz + 1)
    si, sj, K = sl(ix.isc, ix.iec), sl(ix.jsc, ix.jec), slice(0, nz + 1)
    ar = g["area"][si, sj, None]
    area_after = ((ar + xfx_i[si, sj, K] - _sh(xfx_i, 1, 0, si, sj)[:, :, K]) + (ar + yfx_i[si, sj, K] - _sh(yfx_i, 0, 1, si, sj)[:, :, K])) - ar
    new = (zh[si, sj, K] * ar + fx[si, sj, K] - _sh(fx, 1, 0, si, sj)[:, :, K] + fy[si, sj, K] - _sh(fy, 0, 1, si, sj)[:, :, K]) / area_after + (
        gx[si, sj, K] - _sh(gx, 1, 0, si, sj)[:, :, K] + gy[si, sj, K] - _sh(gy, 0, 1, si, sj)[:, :, K]) / ar
    ws[si, sj] = (zs[si, sj] - new[:, :, nz]) / dt
    for k in range(nz - 1, -1, -1):
        other = new[:, :, k + 1] + DZ_MIN
        new[:, :, k] = np.where(new[:, :, k] > other, new[:, :, k], other)
    zh[si, sj, K] = new
