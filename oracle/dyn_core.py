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
