"""TracerAdvection — drop-in for fv3core/pace/fv3core/stencils/tracer_2d_1l.py:166-392."""
import math
from typing import Dict

from ... import constants as c
from ...util.quantity import Quantity

_C3 = (c.X_DIM, c.Y_DIM, c.Z_DIM)


class TracerAdvection:
    def __init__(self, stencil_factory, quantity_factory, transport, grid_data, comm, tracers: Dict[str, Quantity]):
        self._rt = stencil_factory.runtime
        self._tracer_count = len(tracers)
        self.grid_data = grid_data
        qf = quantity_factory
        self._x_area_flux = qf.zeros((c.X_INTERFACE_DIM, c.Y_DIM, c.Z_DIM), "unknown")
        self._y_area_flux = qf.zeros((c.X_DIM, c.Y_INTERFACE_DIM, c.Z_DIM), "unknown")
        self._x_flux = qf.zeros((c.X_INTERFACE_DIM, c.Y_INTERFACE_DIM, c.Z_DIM), "unknown")
        self._y_flux = qf.zeros((c.X_INTERFACE_DIM, c.Y_INTERFACE_DIM, c.Z_DIM), "unknown")
        self._tmp_dp = qf.zeros(_C3, "Pa")
        self.finite_volume_transport = transport
        spec = qf.get_quantity_halo_spec(_C3, n_halo=3)
        self._updater = comm.get_scalar_halo_updater([spec] * self._tracer_count)
        self._tracer_list = list(tracers.values())
        import torch

        self._tracer_ptrs = torch.tensor([q.ptr for q in self._tracer_list], dtype=torch.int64).to(self._rt.device)
        self._fused = abs(transport._hord) == 8

    def __call__(self, tracers: Dict[str, Quantity], dp1: Quantity, x_mass_flux: Quantity, y_mass_flux: Quantity,
                 x_courant: Quantity, y_courant: Quantity):
        rt = self._rt
        # the CFL allreduce is commented out in the reference: cmax = 2 -> 3 sub-cycles (tracer_2d_1l.py:312-339)
        cmax_max_all_ranks = 2.0
        n_split = math.floor(1.0 + cmax_max_all_ranks)
        rt.call("fv3_tracer_flux_prep", x_courant.ptr, y_courant.ptr, x_mass_flux.ptr, y_mass_flux.ptr,
                self._x_area_flux.ptr, self._y_area_flux.ptr, int(n_split))
        qs = list(tracers.values())
        self._updater.update(qs)
        dp2 = self._tmp_dp
        if self._fused and [q.ptr for q in qs] == [q.ptr for q in self._tracer_list]:
            # one launch per sub-cycle for all tracers (fluxes never leave shared memory)
            for it in range(n_split):
                last_call = it == n_split - 1
                rt.call("fv3_tracer_subcycle", self._tracer_ptrs.data_ptr(), len(qs), dp1.ptr, dp2.ptr, x_mass_flux.ptr,
                        y_mass_flux.ptr, x_courant.ptr, y_courant.ptr, self._x_area_flux.ptr, self._y_area_flux.ptr,
                        int(self.finite_volume_transport._hord), 0 if last_call else 1)
                if not last_call:
                    self._updater.update(qs)
            return
        for it in range(n_split):
            last_call = it == n_split - 1
            rt.call("fv3_tracer_apply_mass_flux", dp1.ptr, x_mass_flux.ptr, y_mass_flux.ptr, dp2.ptr)
            for q in qs:
                self.finite_volume_transport(q, x_courant, y_courant, self._x_area_flux, self._y_area_flux, self._x_flux,
                                             self._y_flux, x_mass_flux=x_mass_flux, y_mass_flux=y_mass_flux)
                rt.call("fv3_tracer_apply_flux", q.ptr, dp1.ptr, self._x_flux.ptr, self._y_flux.ptr, dp2.ptr)
            if not last_call:
                self._updater.update(qs)
                rt.call("fv3_tracer_swap_dp", dp1.ptr, dp2.ptr)
