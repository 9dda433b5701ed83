"""`GridData` / `DampingCoefficients`: metric terms of all local subdomains as device Quantities.

Same attribute names as the reference containers (util/pace/util/grid/helper.py:20-640) for every term the
hot path reads.  Instances are built either from this package's own generator
(pace_b200.util.grid.generation) or from per-rank arrays of any other source (`from_arrays`).
"""
from typing import Dict, List, Sequence

import numpy as np
import torch

from ... import constants as c
from ..quantity import Quantity
from ..sizer import QuantityFactory

HORIZONTAL_2D = [
    "dx", "dy", "dxa", "dya", "dxc", "dyc", "rdx", "rdy", "rdxa", "rdya", "rdxc", "rdyc",
    "area", "area_64", "rarea", "rarea_c",
    "cosa", "cosa_u", "cosa_v", "cosa_s", "sina_u", "sina_v", "rsina", "rsin_u", "rsin_v", "rsin2",
    "sin_sg1", "sin_sg2", "sin_sg3", "sin_sg4", "cos_sg1", "cos_sg2", "cos_sg3", "cos_sg4",
    "fC", "fC_agrid", "lon", "lat", "lon_agrid", "lat_agrid",
    "edge_w", "edge_e", "edge_s", "edge_n",
    "a11", "a12", "a21", "a22",
]
DAMPING_2D = ["divg_u", "divg_v", "del6_u", "del6_v"]
COLUMNS = ["ak", "bk", "dp_ref", "p"]
INIT_HOST = ["lon", "lat", "lon_agrid", "lat_agrid", "ee1", "ee2", "es1", "ew2", "ak", "bk", "area"]


def _as_2d(name, arr, shape):
    """edge_s / edge_n are 1-D in i in the reference (FloatFieldI); store them broadcast along j."""
    arr = np.asarray(arr, dtype=np.float64)
    if arr.ndim == 1:
        return np.broadcast_to(arr[:, None], shape).copy()
    return arr


class GridData:
    def __init__(self, fields: Dict[str, Quantity], columns: Dict[str, torch.Tensor], scalars: Dict[str, float]):
        self._fields = fields
        self._columns = columns
        self._scalars = scalars

    @classmethod
    def from_arrays(cls, qf: QuantityFactory, per_rank: Sequence[Dict[str, np.ndarray]]) -> "GridData":
        """per_rank[s][name] = array [i, j] (reference storage order, halo included) of local subdomain s."""
        g = qf.geometry
        dims = (c.X_DIM, c.Y_DIM)
        fields = {}
        for name in HORIZONTAL_2D:
            if name not in per_rank[0]:
                continue
            stack = np.stack([_as_2d(name, r[name], (g.ni, g.nj)) for r in per_rank])
            fields[name] = qf.from_array(stack, dims, "")
        columns = {}
        for name in COLUMNS:
            if name in per_rank[0]:
                col = np.zeros(g.nk)
                a = np.asarray(per_rank[0][name], dtype=np.float64)
                col[: len(a)] = a
                columns[name] = torch.as_tensor(col).to(qf.device)
        scalars = {k: float(per_rank[0][k]) for k in ("ptop", "p_ref") if k in per_rank[0]}
        scalars["ks"] = int(per_rank[0].get("ks", _ks_from_bk(per_rank[0].get("bk"))))
        out = cls(fields, columns, scalars)
        # host copies of what the analytic initial conditions read (positions and grid unit vectors)
        out._host = [{k: np.asarray(r[k]) for k in INIT_HOST if k in r} for r in per_rank]
        return out

    @classmethod
    def new_from_generation(cls, qf: QuantityFactory, comm, nz: int = None) -> "GridData":
        """Metric terms of this process's subdomains from pace_b200.util.grid.generation (the role of
        GridData.new_from_metric_terms(MetricTerms(...)) in the reference, helper.py:60-120)."""
        from . import generation

        nx_tile = comm.decomposition.n_tile
        per_rank = generation.generate(nx_tile, comm.decomposition.layout, comm.local_ranks, nz or qf.geometry.nz)
        out = cls.from_arrays(qf, per_rank)
        out._per_rank = per_rank
        return out

    def host_dicts(self):
        return self._host

    def __getattr__(self, name):
        for store in ("_fields", "_columns", "_scalars"):
            d = self.__dict__.get(store, {})
            if name in d:
                return d[name]
        raise AttributeError(name)

    @property
    def f0(self):
        return self._fields["fC_agrid"]

    def host(self, name) -> np.ndarray:
        v = getattr(self, name)
        return v.numpy() if isinstance(v, Quantity) else v.cpu().numpy()


def _ks_from_bk(bk):
    """ks = number of pure-pressure layers (bk == 0 interfaces - 1); eta.py of the reference tabulates it."""
    if bk is None:
        return 0
    bk = np.asarray(bk)
    nz0 = int(np.sum(bk == 0.0))
    return max(nz0 - 1, 0)


class DampingCoefficients:
    def __init__(self, fields: Dict[str, Quantity], da_min: float, da_min_c: float):
        self._fields = fields
        self.da_min = float(da_min)
        self.da_min_c = float(da_min_c)

    @classmethod
    def from_arrays(cls, qf: QuantityFactory, per_rank: Sequence[Dict[str, np.ndarray]], prefix="damp_"):
        dims = (c.X_DIM, c.Y_DIM)
        fields = {}
        for name in DAMPING_2D:
            fields[name] = qf.from_array(np.stack([np.asarray(r[prefix + name]) for r in per_rank]), dims, "")
        return cls(fields, per_rank[0][prefix + "da_min"], per_rank[0][prefix + "da_min_c"])

    @classmethod
    def new_from_generation(cls, qf: QuantityFactory, grid_data: "GridData"):
        """From the same generated terms as GridData.new_from_generation (DampingCoefficients.new_from_metric_terms
        in the reference, helper.py:20-60)."""
        return cls.from_arrays(qf, grid_data._per_rank)

    def __getattr__(self, name):
        d = self.__dict__.get("_fields", {})
        if name in d:
            return d[name]
        raise AttributeError(name)
