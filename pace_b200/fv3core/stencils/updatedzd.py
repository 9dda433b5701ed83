"""UpdateHeightOnDGrid — drop-in for fv3core/pace/fv3core/stencils/updatedzd.py:199-356."""
import numpy as np
import torch

from ...util.quantity import Quantity


def cubic_spline_interpolation_constants(dp0: np.ndarray):
    """updatedzd.py:129-154 on the host; dp0 = dp_ref[0:nz]."""
    nz = len(dp0)
    gk, beta, gamma = np.zeros(nz), np.zeros(nz), np.zeros(nz)
    gk[0] = dp0[1] / dp0[0]
    beta[0] = gk[0] * (gk[0] + 0.5)
    gamma[0] = (1.0 + gk[0] * (gk[0] + 1.5)) / beta[0]
    gk[1:] = dp0[:-1] / dp0[1:]
    for i in range(1, nz):
        beta[i] = 2.0 + 2.0 * gk[i] - gamma[i - 1]
        gamma[i] = gk[i] / beta[i]
    return gk, beta, gamma


class UpdateHeightOnDGrid:
    def __init__(self, stencil_factory, quantity_factory, damping_coefficients, grid_data, grid_type: int, hord_tm: int,
                 column_namelist):
        self._rt = rt = stencil_factory.runtime
        cols = column_namelist.host if hasattr(column_namelist, "host") else column_namelist
        if any(cols["damp_vt"] <= 1e-5):
            raise NotImplementedError("damp <= 1e-5 in column_namelist is untested")
        if int(hord_tm) != int(rt.config.hord_tm):
            raise NotImplementedError("hord_tm must equal the value in the dycore config")
        nz = rt.comm.geometry.nz
        dp0 = grid_data.host("dp_ref")[:nz]
        self._consts = [self._dev(np.concatenate([c, [0.0]])) for c in cubic_spline_interpolation_constants(dp0)]
        self._damp = self._dev(np.concatenate([cols["damp_vt"], [0.0]]))  # Z_DIM storage has a trailing 0 level
        self._nord = self._dev(np.concatenate([cols["nord_v"], cols["nord_v"][-1:]]))
        self._nmax = int(cols["nord_v"].max())

    def _dev(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self._rt.device)

    def __call__(self, surface_height: Quantity, height: Quantity, courant_number_x: Quantity, courant_number_y: Quantity,
                 x_area_flux: Quantity, y_area_flux: Quantity, ws: Quantity, dt: float):
        gk, beta, gamma = self._consts
        self._rt.call("fv3_update_dz_d", surface_height.ptr, height.ptr, courant_number_x.ptr, courant_number_y.ptr,
                      x_area_flux.ptr, y_area_flux.ptr, ws.ptr, float(dt), gk.data_ptr(), beta.data_ptr(),
                      gamma.data_ptr(), self._damp.data_ptr(), self._nord.data_ptr(), self._nmax)
