"""AdjustNegativeTracerMixingRatio — drop-in for fv3core/pace/fv3core/stencils/neg_adj3.py:316-420."""


class AdjustNegativeTracerMixingRatio:
    """Adjust tracer mixing ratios to fix negative values (neg_adj3 in Fortran): one native call, in place."""

    def __init__(self, stencil_factory, quantity_factory=None, check_negative: bool = False, hydrostatic: bool = False):
        if check_negative:
            raise NotImplementedError("Unimplemented namelist value check_negative=True")
        if hydrostatic:
            raise NotImplementedError("Unimplemented namelist hydrostatic=True")
        self._rt = stencil_factory.runtime

    def __call__(self, qvapor, qliquid, qrain, qsnow, qice, qgraupel, qcld, pt, delp):
        self._rt.call("fv3_neg_adj3", qvapor.ptr, qliquid.ptr, qrain.ptr, qsnow.ptr, qice.ptr, qgraupel.ptr, qcld.ptr,
                      pt.ptr, delp.ptr)
