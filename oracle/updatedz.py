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
