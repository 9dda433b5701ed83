"""UpdateGeopotentialHeightOnCGrid — drop-in for fv3core/pace/fv3core/stencils/updatedzc.py:120-207."""
from ...util.quantity import Quantity
from ..stencil_factory import StencilFactory


class UpdateGeopotentialHeightOnCGrid:
    def __init__(self, stencil_factory: StencilFactory, quantity_factory, area: Quantity, dp_ref):
        self._rt = stencil_factory.runtime  # area and dp_ref are taken from the runtime's grid pointers

    def __call__(self, zs: Quantity, ut: Quantity, vt: Quantity, gz: Quantity, ws: Quantity, dt: float):
        self._rt.call("fv3_update_dz_c", zs.ptr, ut.ptr, vt.ptr, gz.ptr, ws.ptr, float(dt))
