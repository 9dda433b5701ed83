"""CGridShallowWaterDynamics — drop-in for fv3core/pace/fv3core/stencils/c_sw.py:483-766."""
from ... import constants as c
from ...util.quantity import Quantity
from ..stencil_factory import StencilFactory

D3 = (c.X_DIM, c.Y_DIM, c.Z_DIM)


class CGridShallowWaterDynamics:
    def __init__(self, stencil_factory: StencilFactory, quantity_factory, grid_data, nested: bool, grid_type: int,
                 nord: int):
        if grid_type >= 3:
            raise NotImplementedError("unimplemented grid_type >= 3")
        if nested:
            raise NotImplementedError("nested grids are not implemented")
        self._rt = stencil_factory.runtime
        self.grid_data = grid_data
        self.delpc = quantity_factory.zeros(D3, units="unknown")
        self.ptc = quantity_factory.zeros(D3, units="unknown")

    def __call__(self, delp: Quantity, pt: Quantity, u: Quantity, v: Quantity, w: Quantity, uc: Quantity,
                 vc: Quantity, ua: Quantity, va: Quantity, ut: Quantity, vt: Quantity, divgd: Quantity,
                 omga: Quantity, dt2: float):
        self._rt.call("fv3_c_sw", delp.ptr, pt.ptr, u.ptr, v.ptr, w.ptr, uc.ptr, vc.ptr, ua.ptr, va.ptr, ut.ptr,
                      vt.ptr, divgd.ptr, omga.ptr, self.delpc.ptr, self.ptc.ptr, float(dt2))
        return self.delpc, self.ptc
