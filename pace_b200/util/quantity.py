"""`Quantity`: array + dims/origin/extent/units, backed by device memory shared with the CUDA kernels.

Mirrors pace.util.Quantity (util/pace/util/quantity.py:259-565) for what the hot path uses: `.data`, `.view[:]`,
`.dims`, `.units`, `.origin`, `.extent`, `.metadata`, `.np`, `__cuda_array_interface__` (:492-494).
B200-first difference: one process drives ALL subdomains resident on its GPU, so storage carries a leading
subdomain axis — `.data` has shape (n_sub, *shape); `.subdomain(s)` is the reference-shaped per-rank Quantity
sharing the same memory.  Memory order is I-fastest ([s][k][j][i]), the order GT4Py's GPU backends use
(external/gt4py/src/gt4py/storage/layout.py:138-175), with the j stride padded to a multiple of 4 doubles.
"""
import dataclasses
from typing import Tuple

import numpy as np
import torch

from .. import constants


@dataclasses.dataclass
class QuantityMetadata:
    origin: Tuple[int, ...]
    extent: Tuple[int, ...]
    dims: Tuple[str, ...]
    units: str
    data_type: type
    dtype: type
    n_halo: int = constants.N_HALO_DEFAULT


@dataclasses.dataclass
class QuantityHaloSpec:
    """What a halo updater needs to know about a field (util/pace/util/quantity.py:40-60)."""

    n_points: int
    strides: Tuple[int, ...]
    itemsize: int
    shape: Tuple[int, ...]
    origin: Tuple[int, ...]
    extent: Tuple[int, ...]
    dims: Tuple[str, ...]
    numpy_module: object
    dtype: type


class _View:
    """`quantity.view[...]`: indexing relative to the compute-domain origin (quantity.py:70-120)."""

    def __init__(self, q):
        self._q = q

    def _slices(self):
        q = self._q
        return (slice(None),) + tuple(slice(o, o + e) for o, e in zip(q.origin, q.extent))

    def __getitem__(self, index):
        return self._q.data[self._slices()][(slice(None),) + (index if isinstance(index, tuple) else (index,))]

    def __setitem__(self, index, value):
        self._q.data[self._slices()][(slice(None),) + (index if isinstance(index, tuple) else (index,))] = value


class Quantity:
    def __init__(self, data: torch.Tensor, dims, units: str, origin=None, extent=None, batched=True):
        if not isinstance(data, torch.Tensor):
            raise TypeError("pace_b200.Quantity wraps a torch.Tensor (device memory); use QuantityFactory.from_array")
        self._batched = batched
        self._data = data
        nd = data.dim() - (1 if batched else 0)
        if len(dims) != nd:
            raise ValueError(f"received {len(dims)} dimension names for {nd} dimensions: {dims}")
        self._dims = tuple(dims)
        self._units = units
        shape = tuple(data.shape[1:] if batched else data.shape)
        self._origin = tuple(origin) if origin is not None else (0,) * nd
        self._extent = tuple(extent) if extent is not None else tuple(s - o for s, o in zip(shape, self._origin))
        self._view = _View(self) if batched else None

    # -- reference-compatible surface -------------------------------------------------
    @property
    def data(self) -> torch.Tensor:
        return self._data

    @property
    def dims(self):
        return self._dims

    @property
    def units(self):
        return self._units

    @property
    def origin(self):
        return self._origin

    @property
    def extent(self):
        return self._extent

    @property
    def shape(self):
        return tuple(self._data.shape)

    @property
    def view(self):
        return self._view

    @property
    def np(self):
        return np

    @property
    def metadata(self) -> QuantityMetadata:
        return QuantityMetadata(origin=self._origin, extent=self._extent, dims=self._dims, units=self._units,
                                data_type=torch.Tensor, dtype=np.float64)

    @property
    def __cuda_array_interface__(self):
        return self._data.__cuda_array_interface__

    def __dlpack__(self, *a, **k):
        return self._data.__dlpack__(*a, **k)

    # -- B200 extensions ----------------------------------------------------------------
    @property
    def n_sub(self):
        return self._data.shape[0] if self._batched else 1

    @property
    def ptr(self) -> int:
        """Address of element (s=0, i=0, j=0, k=0); what the C ABI takes."""
        return self._data.data_ptr()

    def subdomain(self, s: int) -> "Quantity":
        if not self._batched:
            return self
        return Quantity(self._data[s], self._dims, self._units, self._origin, self._extent, batched=False)

    def numpy(self):
        """Host copy, logical index order (s, i, j, k)."""
        return self._data.detach().cpu().numpy()

    def set_from_numpy(self, arr, s=None):
        t = torch.as_tensor(np.ascontiguousarray(arr), dtype=self._data.dtype)
        if s is None:
            self._data.copy_(t.to(self._data.device))
        else:
            self._data[s].copy_(t.to(self._data.device))

    def __repr__(self):
        return f"Quantity(dims={self._dims}, units={self._units!r}, shape={self.shape}, origin={self._origin}, extent={self._extent})"
