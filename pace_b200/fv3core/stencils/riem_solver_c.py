"""NonhydrostaticVerticalSolverCGrid — drop-in for fv3core/pace/fv3core/stencils/riem_solver_c.py:126-250."""
from ...util.quantity import Quantity
from ..stencil_factory import StencilFactory


class NonhydrostaticVerticalSolverCGrid:
    def __init__(self, stencil_factory: StencilFactory, quantity_factory, p_fac: float):
        self._rt = stencil_factory.runtime
        if abs(p_fac - self._rt.config.p_fac) > 0:
            raise NotImplementedError("p_fac must equal the value in the dycore config")

    def __call__(self, dt2: float, cappa: Quantity, ptop: float, hs: Quantity, ws: Quantity, ptc: Quantity,
                 q_con: Quantity, delpc: Quantity, gz: Quantity, pef: Quantity, w3: Quantity):
        self._rt.call("fv3_riem_solver_c", float(dt2), cappa.ptr, float(ptop), hs.ptr, ws.ptr, ptc.ptr, q_con.ptr,
                      delpc.ptr, gz.ptr, pef.ptr, w3.ptr)
