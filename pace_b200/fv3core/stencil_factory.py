"""`StencilFactory` / `GridIndexing` stand-ins.

The reference's StencilFactory compiles GT4Py stencils (dsl/pace/dsl/stencil.py:858-1002); that layer is deleted
here.  The object is kept as the handle stage constructors receive so their signatures stay those of the
reference; it carries `.grid_indexing` (domain sizes and tile-edge flags, stencil.py:542-855) and the native
`Runtime` the CUDA entry points are called through.
"""
import dataclasses
from typing import Tuple

from .runtime import Runtime


@dataclasses.dataclass
class GridIndexing:
    domain: Tuple[int, int, int]
    n_halo: int
    edge_flags: Tuple[Tuple[bool, bool, bool, bool], ...]  # per local subdomain: (south, north, west, east)

    @property
    def isc(self):
        return self.n_halo

    @property
    def iec(self):
        return self.n_halo + self.domain[0] - 1

    @property
    def jsc(self):
        return self.n_halo

    @property
    def jec(self):
        return self.n_halo + self.domain[1] - 1

    @classmethod
    def from_sizer_and_communicator(cls, sizer, cube):
        flags = []
        for r in cube.local_ranks:
            w, e, s, n = cube.decomposition.edge_flags(r)
            flags.append((s, n, w, e))
        return cls((sizer.nx, sizer.ny, sizer.nz), sizer.n_halo, tuple(flags))


class StencilFactory:
    def __init__(self, config=None, grid_indexing: GridIndexing = None, runtime: Runtime = None):
        self.config = config
        self.grid_indexing = grid_indexing
        self.runtime = runtime
