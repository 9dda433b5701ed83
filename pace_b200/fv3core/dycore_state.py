"""`DycoreState`: the prognostic/diagnostic fields of the dycore as device Quantities.

Field names, dims and units follow fv3core/pace/fv3core/initialization/dycore_state.py:10-282.
"""
import dataclasses
from typing import Dict

import numpy as np

from .. import constants as c
from ..util.quantity import Quantity

_C = (c.X_DIM, c.Y_DIM, c.Z_DIM)
_CI = (c.X_DIM, c.Y_DIM, c.Z_INTERFACE_DIM)
_U = (c.X_DIM, c.Y_INTERFACE_DIM, c.Z_DIM)
_V = (c.X_INTERFACE_DIM, c.Y_DIM, c.Z_DIM)
_2D = (c.X_DIM, c.Y_DIM)

FIELDS = {
    "u": (_U, "m/s"), "v": (_V, "m/s"), "w": (_C, "m/s"), "ua": (_C, "m/s"), "va": (_C, "m/s"), "uc": (_V, "m/s"),
    "vc": (_U, "m/s"), "delp": (_C, "Pa"), "delz": (_C, "m"), "ps": (_2D, "Pa"), "pe": (_CI, "Pa"), "pt": (_C, "degK"),
    "peln": (_CI, "ln(Pa)"), "pk": (_CI, "unknown"), "pkz": (_C, "unknown"), "qvapor": (_C, "kg/kg"),
    "qliquid": (_C, "kg/kg"), "qice": (_C, "kg/kg"), "qrain": (_C, "kg/kg"), "qsnow": (_C, "kg/kg"),
    "qgraupel": (_C, "kg/kg"), "qo3mr": (_C, "kg/kg"), "qsgs_tke": (_C, "m**2/s**2"), "qcld": (_C, ""),
    "q_con": (_C, "kg/kg"), "omga": (_C, "Pa/s"), "mfxd": (_V, "unknown"), "mfyd": (_U, "unknown"), "cxd": (_V, ""),
    "cyd": (_U, ""), "diss_estd": (_C, "unknown"), "phis": (_2D, "m^2 s^-2"),
}
# dsl/pace/dsl/gt4py_utils.py:24-34
TRACER_VARIABLES = ["qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke", "qcld"]


class DycoreState:
    def __init__(self, **quantities):
        for name, (dims, units) in FIELDS.items():
            q = quantities[name]
            if tuple(q.dims) != tuple(dims):
                raise TypeError(f"{name} has metadata dims of {q.dims} that does not match the requirement {dims}")
            setattr(self, name, q)
        self.bdt = quantities.get("bdt", 0.0)
        self.mdt = quantities.get("mdt", 0.0)

    @classmethod
    def init_zeros(cls, quantity_factory) -> "DycoreState":
        return cls(**{n: quantity_factory.zeros(d, u) for n, (d, u) in FIELDS.items()})

    @classmethod
    def init_from_numpy_arrays(cls, per_rank: list, quantity_factory) -> "DycoreState":
        """per_rank[s][name]: array [i, j(, k)] of local subdomain s (reference storage order)."""
        for name in per_rank[0]:
            if name not in FIELDS:
                raise KeyError(name + " is provided, but not part of the dycore state")
        out = {}
        for name, (dims, units) in FIELDS.items():
            out[name] = quantity_factory.from_array(np.stack([np.asarray(r[name], dtype=np.float64) for r in per_rank]), dims, units)
        return cls(**out)

    def as_numpy(self, s: int = None) -> Dict[str, np.ndarray]:
        return {n: (getattr(self, n).numpy() if s is None else getattr(self, n).numpy()[s]) for n in FIELDS}

    @property
    def tracers(self) -> Dict[str, Quantity]:
        return {n: getattr(self, n) for n in TRACER_VARIABLES[:8]}
