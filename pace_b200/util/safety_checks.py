"""State sanity checks with the reference driver's interface, evaluated on the device.

Mirrors `pace.driver.safety_checks` (driver/pace/driver/safety_checks.py:8-110): `VariableBounds`,
`SafetyChecker.register_variable / clear_all_checks / check_state` with the same semantics and the same exceptions
(NotImplementedError for a doubly registered or unknown variable, RuntimeError for a value outside its bounds or a NaN).
The min / max / NaN scan of a field is ONE kernel pass (`fv3_field_check`, csrc/check.cu) instead of three numpy
reductions over a host copy; only three 64-bit words come back.
"""
import dataclasses
from typing import ClassVar, Dict, Optional

import numpy as np
import torch

_MASK = 0x7FFFFFFFFFFFFFFF


def _value_of(key: int) -> float:
    """Inverse of the order-preserving key of csrc/check.cu (the map is an involution on the bit pattern)."""
    bits = key ^ ((key >> 63) & _MASK)
    return float(np.array([bits], dtype=np.int64).view(np.float64)[0])


def field_min_max_nan(rt, q, compute_domain_only=False):
    """(min, max, number of NaNs) of a Quantity over its view (origin / extent: `q.view[:]` of the reference) or over its
    whole storage (`q.data`); min / max ignore NaNs and are (+inf, -inf) when every value is a NaN."""
    three_d = len(q.dims) == 3
    shape = q.shape[1:] if q.n_sub > 1 or len(q.shape) == len(q.dims) + 1 else q.shape   # logical (i, j[, k]) extents
    if compute_domain_only:
        i0, j0 = q.origin[0], q.origin[1]
        i1, j1 = i0 + q.extent[0], j0 + q.extent[1]
        nk = q.origin[2] + q.extent[2] if three_d else 0
        if three_d and q.origin[2] != 0:
            raise NotImplementedError("fv3_field_check scans levels from 0")
    else:
        i0, j0, i1, j1 = 0, 0, shape[0], shape[1]
        nk = shape[2] if three_d else 0
    out = torch.empty(3, dtype=torch.int64, device=q.data.device)
    rt.call("fv3_field_check", q.ptr, int(i0), int(i1), int(j0), int(j1), int(nk), out.data_ptr())
    kmin, kmax, nnan = (int(v) for v in out.cpu().tolist())
    vmin = float("inf") if kmin == np.iinfo(np.int64).max else _value_of(kmin)
    vmax = float("-inf") if kmax == np.iinfo(np.int64).min else _value_of(kmax)
    return vmin, vmax, nnan


@dataclasses.dataclass
class VariableBounds:
    minimum_value: Optional[float] = None
    maximum_value: Optional[float] = None
    compute_domain_only: bool = False


class SafetyChecker:
    """Checks the registered variables of a DycoreState against their bounds and for NaNs (safety_checks.py:24-110)."""

    checks: ClassVar[Dict[str, VariableBounds]] = {}

    def __init__(self, runtime):
        self.rt = runtime

    @classmethod
    def register_variable(cls, name: str, minimum_value: Optional[float] = None, maximum_value: Optional[float] = None,
                          compute_domain_only: bool = False):
        if name in cls.checks:
            raise NotImplementedError("Can only register variables once")
        cls.checks[name] = VariableBounds(minimum_value, maximum_value, compute_domain_only)

    @classmethod
    def clear_all_checks(cls):
        cls.checks.clear()

    def check_state(self, state):
        for variable, bounds in self.checks.items():
            try:
                var = getattr(state, variable)
            except AttributeError:
                raise NotImplementedError("Variable is not in the state")
            vmin, vmax, nnan = field_min_max_nan(self.rt, var, bounds.compute_domain_only)
            # the reference tests truthiness of the bound (a bound of 0.0 is "no bound"): kept
            if bounds.minimum_value and vmin < bounds.minimum_value:
                raise RuntimeError(f"Variable {variable} is outside of its specified bounds: "
                                   f"{bounds.minimum_value} specified, {vmin} found")
            if bounds.maximum_value and vmax > bounds.maximum_value:
                raise RuntimeError(f"Variable {variable} is outside of its specified bounds: "
                                   f"{bounds.maximum_value} specified, {vmax} found")
            # the NaN test of the reference looks at the compute domain (`var.view[:]`) whatever the bounds looked at
            if not bounds.compute_domain_only:
                nnan = field_min_max_nan(self.rt, var, True)[2]
            if nnan:
                raise RuntimeError(f"Variable {variable} contains a NaN value")
