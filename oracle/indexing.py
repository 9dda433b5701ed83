"""Index bookkeeping of one subdomain for the oracle (GridIndexing semantics, dsl/pace/dsl/stencil.py:542-855)."""
import dataclasses


@dataclasses.dataclass
class Idx:
    nx: int
    ny: int
    nz: int
    west: bool = True
    east: bool = True
    south: bool = True
    north: bool = True
    halo: int = 3

    @property
    def isc(self):
        return self.halo

    @property
    def iec(self):
        return self.halo + self.nx - 1

    @property
    def jsc(self):
        return self.halo

    @property
    def jec(self):
        return self.halo + self.ny - 1

    isd = jsd = 0

    @property
    def ied(self):
        return self.iec + self.halo

    @property
    def jed(self):
        return self.jec + self.halo

    @classmethod
    def from_edge_mask(cls, nx, ny, nz, mask):
        return cls(nx, ny, nz, bool(mask & 1), bool(mask & 2), bool(mask & 4), bool(mask & 8))


def sl(a, b):
    """inclusive index range -> slice"""
    return slice(a, b + 1)
