"""Grid sizing, storage geometry and allocation.

`SubtileGridSizer` / `QuantityFactory` mirror util/pace/util/initialization/sizer.py:33-155 and
allocator.py:31-155: every field is allocated with horizontal shape n+1+2*halo and vertical shape nz+1.
`Geometry` adds what the CUDA side needs: padded strides and the per-subdomain tile-edge flags.
"""
import dataclasses
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from .. import constants as c
from .quantity import Quantity, QuantityHaloSpec


@dataclasses.dataclass
class SubtileGridSizer:
    nx: int
    ny: int
    nz: int
    n_halo: int = c.N_HALO_DEFAULT
    extra_dim_lengths: Dict[str, int] = dataclasses.field(default_factory=dict)

    @classmethod
    def from_tile_params(cls, nx_tile, ny_tile, nz, n_halo=c.N_HALO_DEFAULT, extra_dim_lengths=None, layout=(1, 1),
                         tile_partitioner=None, tile_rank=0):
        if nx_tile % layout[1] or ny_tile % layout[0]:
            raise ValueError("tile size must be divisible by the layout")
        return cls(nx_tile // layout[1], ny_tile // layout[0], nz, n_halo, dict(extra_dim_lengths or {}))

    @property
    def dim_extents(self):
        return {c.X_DIM: self.nx, c.X_INTERFACE_DIM: self.nx + 1, c.Y_DIM: self.ny, c.Y_INTERFACE_DIM: self.ny + 1,
                c.Z_DIM: self.nz, c.Z_INTERFACE_DIM: self.nz + 1, **self.extra_dim_lengths}

    def get_origin(self, dims: Sequence[str]) -> Tuple[int, ...]:
        return tuple(self.n_halo if d in c.HORIZONTAL_DIMS else 0 for d in dims)

    def get_extent(self, dims: Sequence[str]) -> Tuple[int, ...]:
        ext = self.dim_extents
        return tuple(ext[d] for d in dims)

    def get_shape(self, dims: Sequence[str]) -> Tuple[int, ...]:
        out = []
        for d in dims:
            if d in c.X_DIMS:
                out.append(self.nx + 1 + 2 * self.n_halo)
            elif d in c.Y_DIMS:
                out.append(self.ny + 1 + 2 * self.n_halo)
            elif d in c.Z_DIMS:
                out.append(self.nz + 1)
            else:
                out.append(self.extra_dim_lengths[d])
        return tuple(out)


@dataclasses.dataclass
class Geometry:
    """Storage geometry shared by Python and the C ABI (struct fv3_geom, include/fv3_b200.h)."""

    n_sub: int
    nx: int
    ny: int
    nz: int
    halo: int
    edge: Tuple[int, ...]  # FV3_EDGE_* bit mask per local subdomain

    @property
    def ni(self):
        return self.nx + 2 * self.halo + 1

    @property
    def nj(self):
        return self.ny + 2 * self.halo + 1

    @property
    def nk(self):
        return self.nz + 1

    @property
    def sj(self):
        return (self.ni + 3) // 4 * 4

    @property
    def sk(self):
        return self.sj * self.nj

    @property
    def ss(self):
        return self.sk * self.nk

    @property
    def ss2(self):
        return self.sk

    def to_c(self):
        from .._lib import Geom

        g = Geom()
        for name in ("n_sub", "nx", "ny", "nz", "halo", "ni", "nj", "nk", "sj", "sk", "ss", "ss2"):
            setattr(g, name, getattr(self, name))
        for s, e in enumerate(self.edge):
            g.edge[s] = e
        return g


class QuantityFactory:
    """Allocates batched, I-fastest device storage (allocator.py:58-155 API: zeros/empty/from_array/...)."""

    def __init__(self, sizer: SubtileGridSizer, geometry: Geometry, device="cuda"):
        self.sizer = sizer
        self.geometry = geometry
        self.device = torch.device(device)

    @classmethod
    def from_backend(cls, sizer, backend: str = "b200", geometry: Geometry = None, device="cuda"):
        if geometry is None:
            geometry = Geometry(1, sizer.nx, sizer.ny, sizer.nz, sizer.n_halo, (15,))
        return cls(sizer, geometry, device)

    def _alloc(self, dims, fill):
        g = self.geometry
        kinds = tuple("x" if d in c.X_DIMS else "y" if d in c.Y_DIMS else "z" if d in c.Z_DIMS else "?" for d in dims)
        if kinds == ("x", "y", "z"):
            base = torch.full((g.n_sub, g.nk, g.nj, g.sj), fill, dtype=torch.float64, device=self.device)
            return base.permute(0, 3, 2, 1)[:, : g.ni]
        if kinds == ("x", "y"):
            base = torch.full((g.n_sub, g.nj, g.sj), fill, dtype=torch.float64, device=self.device)
            return base.permute(0, 2, 1)[:, : g.ni]
        if kinds == ("z",):
            return torch.full((g.n_sub, g.nk), fill, dtype=torch.float64, device=self.device)
        raise NotImplementedError(f"unsupported dims {dims}")

    def zeros(self, dims, units, dtype="float64") -> Quantity:
        return Quantity(self._alloc(dims, 0.0), dims, units, self.sizer.get_origin(dims), self.sizer.get_extent(dims))

    def empty(self, dims, units, dtype="float64") -> Quantity:
        return self.zeros(dims, units, dtype)

    def ones(self, dims, units, dtype="float64") -> Quantity:
        return Quantity(self._alloc(dims, 1.0), dims, units, self.sizer.get_origin(dims), self.sizer.get_extent(dims))

    def from_array(self, data, dims, units) -> Quantity:
        """data: array of logical shape (n_sub, *shape(dims)) or (*shape(dims)) when n_sub == 1."""
        q = self.zeros(dims, units)
        arr = np.asarray(data)
        if arr.ndim == len(dims):
            arr = arr[None]
        q.set_from_numpy(arr)
        return q

    def get_quantity_halo_spec(self, dims, n_halo=None, dtype="float64") -> QuantityHaloSpec:
        shape = self.sizer.get_shape(dims)
        g = self.geometry
        strides = {"x": 8, "y": 8 * g.sj, "z": 8 * g.sk}
        st = tuple(strides["x" if d in c.X_DIMS else "y" if d in c.Y_DIMS else "z"] for d in dims)
        return QuantityHaloSpec(
            n_points=self.sizer.n_halo if n_halo is None else n_halo, strides=st, itemsize=8, shape=shape,
            origin=self.sizer.get_origin(dims), extent=self.sizer.get_extent(dims), dims=tuple(dims),
            numpy_module=np, dtype=np.float64,
        )
