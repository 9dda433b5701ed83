"""FiniteVolumeTransport / DelnFluxNoSG — drop-ins for fvtp2d.py:96-346 and delnflux.py:1061-1261 of the reference."""
import numpy as np
import torch

from ...util.quantity import Quantity
from ..stencil_factory import StencilFactory


def calc_damp(damp_c: np.ndarray, da_min: float, nord: np.ndarray) -> np.ndarray:
    """delnflux.calc_damp (delnflux.py:18-33): (damp_c * da_min) ** (nord + 1), per level, on the host."""
    return (np.asarray(damp_c, dtype=np.float64) * da_min) ** (np.asarray(nord, dtype=np.float64) + 1)


def _column(rt, values) -> torch.Tensor:
    col = np.zeros(rt.comm.geometry.nk + 1)
    v = np.asarray(values, dtype=np.float64)
    col[: len(v)] = v
    if len(v) < len(col):
        col[len(v):] = v[-1]  # interface fields (nz+1 levels) reuse the last layer's coefficient
    return torch.as_tensor(col).to(rt.device)


class FiniteVolumeTransport:
    def __init__(self, stencil_factory: StencilFactory, quantity_factory, grid_data, damping_coefficients,
                 grid_type: int, hord: int, nord=None, damp_c=None):
        if grid_type >= 3:
            raise NotImplementedError("grid_type >= 3 is not implemented")
        if abs(hord) not in (5, 6, 8):
            raise NotImplementedError("only hord 5, 6 and 8 are implemented (xppm.py:161,256)")
        self._rt = stencil_factory.runtime
        self._hord = int(hord)
        self._nord = self._damp = None
        self._nmax = 0
        self._no_compute = True
        if nord is not None and damp_c is not None:
            nord = np.asarray(nord, dtype=np.float64)
            damp_c = np.asarray(damp_c, dtype=np.float64)
            if not (damp_c <= 1e-4).all():
                if (damp_c[:-1] <= 1e-4).any():
                    raise NotImplementedError("damp_c currently must be always greater than 10^-4 for delnflux")
                if not all(n in (0, 2, 3) for n in nord):
                    raise NotImplementedError("nord must have values 0, 2, or 3")
                self._no_compute = False
                self._nmax = int(nord.max())
                self._nord = _column(self._rt, nord)
                self._damp = _column(self._rt, calc_damp(damp_c, damping_coefficients.da_min, nord))

    def __call__(self, q: Quantity, crx: Quantity, cry: Quantity, x_area_flux: Quantity, y_area_flux: Quantity,
                 q_x_flux: Quantity, q_y_flux: Quantity, x_mass_flux: Quantity = None, y_mass_flux: Quantity = None,
                 mass: Quantity = None, nk: int = None):
        g = self._rt.comm.geometry
        nk = g.nz if nk is None else nk
        nord = None if self._no_compute else self._nord.data_ptr()
        damp = None if self._no_compute else self._damp.data_ptr()
        self._rt.call("fv3_fvtp2d", q.ptr, crx.ptr, cry.ptr, x_area_flux.ptr, y_area_flux.ptr, q_x_flux.ptr,
                      q_y_flux.ptr, x_mass_flux.ptr if x_mass_flux is not None else None,
                      y_mass_flux.ptr if y_mass_flux is not None else None, mass.ptr if mass is not None else None,
                      self._hord, nord, damp, self._nmax, nk)


class DelnFluxNoSG:
    def __init__(self, stencil_factory: StencilFactory, damping_coefficients, rarea, nord, nk: int = None):
        self._rt = stencil_factory.runtime
        nord = np.asarray(nord, dtype=np.float64)
        self._nmax = int(nord.max())
        if self._nmax > 3:
            raise ValueError("nord must be less than 3")
        if not all(n in (0, 2, 3) for n in nord):
            raise NotImplementedError("nord must have values 0, 2, or 3")
        self._nord = _column(self._rt, nord)
        self._nk = nk

    def __call__(self, q: Quantity, fx2: Quantity, fy2: Quantity, damp_c: torch.Tensor, d2: Quantity = None, mass=None):
        if mass is not None:
            raise NotImplementedError("DelnFluxNoSG with mass is reached through FiniteVolumeTransport only")
        nk = self._nk or self._rt.comm.geometry.nz
        self._rt.call("fv3_delnflux_nosg", q.ptr, fx2.ptr, fy2.ptr, damp_c.data_ptr(), self._nord.data_ptr(),
                      self._nmax, nk)
