"""Drop-ins for the remaining per-substep stage classes of the reference:
NonhydrostaticVerticalSolver (riem_solver3.py:148-321), PK3Halo (pk3_halo.py:41-69),
NonHydrostaticPressureGradient (nh_p_grad.py:115-255), RayleighDamping (ray_fast.py:155-206),
HyperdiffusionDamping (del2cubed.py:68-194)."""
import math

import numpy as np
import torch

from ... import constants
from ...util.quantity import Quantity

SDAY = 86400.0


class NonhydrostaticVerticalSolver:
    def __init__(self, stencil_factory, quantity_factory, config):
        if config.a_imp <= 0.999:
            raise NotImplementedError("a_imp <= 0.999 is not implemented")
        if config.use_logp or config.beta != 0:
            raise NotImplementedError("use_logp / beta != 0 are not implemented")
        self._rt = stencil_factory.runtime

    def __call__(self, last_call: bool, dt: float, cappa: Quantity, ptop: float, zs: Quantity, ws: Quantity,
                 delz: Quantity, q_con: Quantity, delp: Quantity, pt: Quantity, zh: Quantity, p: Quantity, ppe: Quantity,
                 pk3: Quantity, pk: Quantity, log_p_interface: Quantity, w: Quantity):
        self._rt.call("fv3_riem_solver3", int(bool(last_call)), float(dt), cappa.ptr, float(ptop), zs.ptr, ws.ptr,
                      delz.ptr, q_con.ptr, delp.ptr, pt.ptr, zh.ptr, p.ptr, ppe.ptr, pk3.ptr, pk.ptr,
                      log_p_interface.ptr, w.ptr)


class PK3Halo:
    def __init__(self, stencil_factory, quantity_factory):
        self._rt = stencil_factory.runtime

    def __call__(self, pk3: Quantity, delp: Quantity, ptop: float, akap: float):
        self._rt.call("fv3_pk3_halo", pk3.ptr, delp.ptr, float(ptop), float(akap))


class NonHydrostaticPressureGradient:
    def __init__(self, stencil_factory, quantity_factory, grid_data, grid_type):
        if grid_type >= 3:
            raise NotImplementedError("grid_type >= 3 is not implemented")
        self._rt = stencil_factory.runtime

    def __call__(self, u: Quantity, v: Quantity, pp: Quantity, gz: Quantity, pk3: Quantity, delp: Quantity, dt: float,
                 ptop: float, akap: float):
        self._rt.call("fv3_nh_p_grad", u.ptr, v.ptr, pp.ptr, gz.ptr, pk3.ptr, delp.ptr, float(dt), float(ptop),
                      float(akap))


class RayleighDamping:
    def __init__(self, stencil_factory, rf_cutoff, tau, hydrostatic):
        self._rt = stencil_factory.runtime
        self._rf_cutoff = float(rf_cutoff)
        self._tau = float(tau)
        self._cache = {}

    def _columns(self, dt, ptop):
        key = (float(dt), float(ptop))
        if key not in self._cache:
            gd = self._rt.grid_data
            nz = self._rt.comm.geometry.nz
            pfull = gd.host("p")[:nz]
            dp = gd.host("dp_ref")[:nz]
            nudge = self._rf_cutoff + min(100.0, 10.0 * ptop)
            tau0 = self._tau * SDAY
            rf = np.ones(nz + 1)
            n_rf = n_nudge = 0
            p_ref = 0.0
            with np.errstate(all="ignore"):
                for k in range(nz):
                    if pfull[k] < self._rf_cutoff:
                        # compute_rf_vals / compute_rff_vals (ray_fast.py:22-38), numpy scalar arithmetic
                        rfv = dt / tau0 * np.sin(0.5 * constants.PI * np.log(self._rf_cutoff / pfull[k])
                                                 / np.log(self._rf_cutoff / ptop)) ** 2
                        rf[k] = 1.0 / (1.0 + rfv)
                        n_rf = k + 1
                    if pfull[k] < nudge:
                        p_ref = dp[k] if k == 0 else p_ref + dp[k]
                        n_nudge = k + 1
            self._cache[key] = (torch.as_tensor(rf).to(self._rt.device), n_rf, n_nudge, float(p_ref))
        return self._cache[key]

    def __call__(self, u: Quantity, v: Quantity, w: Quantity, dp=None, pfull=None, dt: float = 0.0, ptop: float = 0.0):
        rf, n_rf, n_nudge, p_ref = self._columns(dt, ptop)
        if n_rf == 0:
            return
        self._rt.call("fv3_ray_fast", u.ptr, v.ptr, w.ptr, rf.data_ptr(), n_rf, n_nudge, p_ref)


class HyperdiffusionDamping:
    def __init__(self, stencil_factory, quantity_factory, damping_coefficients, rarea, nmax: int):
        self._rt = stencil_factory.runtime
        self._nmax = int(nmax)

    def __call__(self, qdel: Quantity, cd: float):
        self._rt.call("fv3_del2cubed", qdel.ptr, float(cd), self._nmax, self._rt.comm.geometry.nz)
