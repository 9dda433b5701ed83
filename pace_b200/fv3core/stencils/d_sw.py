"""D-grid shallow-water Lagrangian dynamics — drop-in for fv3core/pace/fv3core/stencils/d_sw.py."""
import ctypes
from typing import Dict

import numpy as np

from ... import _lib
from ...util.quantity import Quantity
from .fvtp2d import _column, calc_damp


def get_column_namelist(config, npz: int) -> Dict[str, np.ndarray]:
    """Per-level damping parameters (d_sw.py:611-683), as host arrays of length npz."""
    col = {}
    for name in ("ke_bg", "d_con", "nord"):
        col[name] = np.full(npz, float(getattr(config, name)))
    col["d2_divg"] = np.full(npz, min(0.2, config.d2_bg))
    col["nord_v"] = np.full(npz, min(2.0, col["nord"][0]))
    col["nord_w"] = np.full(npz, col["nord_v"][0])
    col["nord_t"] = np.full(npz, col["nord_v"][0])
    col["damp_vt"] = np.full(npz, float(config.vtdm4) if config.do_vort_damp else 0.0)
    col["damp_w"] = np.full(npz, col["damp_vt"][0])
    col["damp_t"] = np.full(npz, col["damp_vt"][0])

    def set_low_kvals(k):
        for name in ("nord", "nord_w", "d_con"):
            col[name][k] = 0
        col["damp_w"][k] = col["d2_divg"][k]

    def lowest_kvals(k):
        set_low_kvals(k)
        if config.do_vort_damp:
            col["nord_v"][k] = 0
            col["damp_vt"][k] = 0.5 * col["d2_divg"][k]

    if npz == 1 or config.n_sponge < 0:
        col["d2_divg"][0] = config.d2_bg
    else:
        col["d2_divg"][0] = max(0.01, config.d2_bg, config.d2_bg_k1)
        lowest_kvals(0)
        if config.d2_bg_k2 > 0.01:
            col["d2_divg"][1] = max(config.d2_bg, config.d2_bg_k2)
            lowest_kvals(1)
        if config.d2_bg_k2 > 0.05:
            col["d2_divg"][2] = max(config.d2_bg, 0.2 * config.d2_bg_k2)
            set_low_kvals(2)
    return col


class ColumnNamelist:
    """Device copies of the per-level damping columns + the native struct fv3_dsw_cols."""

    def __init__(self, rt, config, damping_coefficients):
        npz = rt.comm.geometry.nz
        self.host = get_column_namelist(config, npz)
        h = self.host
        derived = {
            "dn_damp_vt": calc_damp(h["damp_vt"], damping_coefficients.da_min, h["nord_v"]),
            "dn_damp_t": calc_damp(h["damp_t"], damping_coefficients.da_min, h["nord_t"]),
            "dn_damp_vt_c": calc_damp(h["damp_vt"], damping_coefficients.da_min_c, h["nord_v"]),
            "dn_damp_w_c": calc_damp(h["damp_w"], damping_coefficients.da_min_c, h["nord_w"]),
        }
        self.device = {}
        self.c = _lib.DswCols()
        for name in _lib.DSW_COLS_PTR:
            t = _column(rt, derived[name] if name in derived else h[name])
            self.device[name] = t
            setattr(self.c, name, t.data_ptr())
        self.c.nmax_v = int(h["nord_v"].max())
        self.c.nmax_w = int(h["nord_w"].max())
        self.c.nmax_t = int(h["nord_t"].max())
        nz_k, nz_nord = 0, int(config.nord)
        for k in range(npz):
            if h["nord"][k] > 0:
                nz_k, nz_nord = k, int(h["nord"][k])
                break
        self.c.nonzero_nord_k = nz_k
        self.c.nonzero_nord = nz_nord

    @property
    def ref(self):
        return ctypes.byref(self.c)


class DGridShallowWaterLagrangianDynamics:
    """Drop-in for d_sw.py:686-1237 (constructor keeps the reference argument list)."""

    def __init__(self, stencil_factory, quantity_factory, grid_data, damping_coefficients, column_namelist,
                 nested: bool, stretched_grid: bool, config):
        if config.grid_type >= 3:
            raise NotImplementedError("ubke and vbke only implemented for grid_type < 3")
        if config.inline_q:
            raise NotImplementedError("inline_q not yet implemented")
        if config.d_ext > 0:
            raise NotImplementedError("untested d_ext > 0. need to call a2b_ord2, not yet implemented")
        if nested or stretched_grid:
            raise NotImplementedError("nested / stretched grids are not implemented")
        if config.do_f3d:
            raise NotImplementedError("do_f3d is not implemented")
        self._rt = stencil_factory.runtime
        self._cols = column_namelist if isinstance(column_namelist, ColumnNamelist) else ColumnNamelist(
            self._rt, config, damping_coefficients)
        h = self._cols.host
        if not ((h["damp_vt"] > 1e-5).all() and (h["damp_w"] > 1e-5).all()):
            raise NotImplementedError("damp_vt and damp_w must exceed 1e-5 on every level (d_sw.py:757-758)")

    def __call__(self, delpc: Quantity, delp, pt, u, v, w, uc, vc, ua, va, divgd, mfx, mfy, cx, cy, crx, cry, xfx, yfx,
                 q_con, zh, heat_source, diss_est, dt: float):
        self._rt.call("fv3_d_sw", delpc.ptr, delp.ptr, pt.ptr, u.ptr, v.ptr, w.ptr, uc.ptr, vc.ptr, ua.ptr, va.ptr,
                      divgd.ptr, mfx.ptr, mfy.ptr, cx.ptr, cy.ptr, crx.ptr, cry.ptr, xfx.ptr, yfx.ptr, q_con.ptr,
                      zh.ptr, heat_source.ptr, diss_est.ptr, float(dt), self._cols.ref)
