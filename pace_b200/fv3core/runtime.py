"""Native context shared by all stage objects of one process: geometry, config, metric-term pointers, scratch."""
import ctypes

import numpy as np

import torch

from .. import _lib
from ..util.grid.helper import DampingCoefficients, GridData


def _great_circle_dist(p1a, p1b, p2a, p2b):
    """a2b_ord4.great_circle_dist (a2b_ord4.py:37-41), evaluated with numpy on the host."""
    tb = np.sin((p1b - p2b) / 2.0) ** 2.0
    ta = np.sin((p1a - p2a) / 2.0) ** 2.0
    return np.arcsin(np.sqrt(tb + np.cos(p1b) * np.cos(p2b) * ta)) * 2.0


def a2b_corner_weights(grid_data, g) -> np.ndarray:
    """Weights x1 / (x2 - x1) of extrap_corner (a2b_ord4.py:43-56) for the 4 corner points of every local
    subdomain and the 3 arms (into the tile, across the x edge, across the y edge)."""
    lon, lat = grid_data.host("lon"), grid_data.host("lat")
    lona, lata = grid_data.host("lon_agrid"), grid_data.host("lat_agrid")
    h, nx, ny = g.halo, g.nx, g.ny
    out = np.zeros((g.n_sub, 4, 3))
    for s in range(g.n_sub):
        for c in range(4):
            west, south = (c % 2 == 0), (c < 2)
            i = h if west else h + nx
            j = h if south else h + ny
            di, dj = (1 if west else -1), (1 if south else -1)
            i0, j0 = (i if west else i - 1), (j if south else j - 1)
            arms = [((i0, j0), (i0 + di, j0 + dj)), ((i0 - di, j0), (i0 - 2 * di, j0 + dj)),
                    ((i0, j0 - dj), (i0 + di, j0 - 2 * dj))]
            for a, (p1, p2) in enumerate(arms):
                x1 = _great_circle_dist(lona[s][p1], lata[s][p1], lon[s, i, j], lat[s, i, j])
                x2 = _great_circle_dist(lona[s][p2], lata[s][p2], lon[s, i, j], lat[s, i, j])
                out[s, c, a] = x1 / (x2 - x1)
    return out


class Runtime:
    def __init__(self, comm, quantity_factory, grid_data: GridData, damping: DampingCoefficients, config):
        self.lib = _lib.load()
        self.comm = comm
        self.qf = quantity_factory
        self.grid_data = grid_data
        self.damping = damping
        self.config = config
        self.device = quantity_factory.device
        if self.device.type != "cuda" and not self.lib.fv3_is_hostsim():
            raise RuntimeError("pace_b200 runs on CUDA devices only (no CPU fallback)")
        g = comm.geometry
        cfg = _lib.Config()
        for name in _lib.CONFIG_INT:
            if hasattr(config, name):
                setattr(cfg, name, int(getattr(config, name)))
        for name in _lib.CONFIG_DBL:
            if hasattr(config, name):
                setattr(cfg, name, float(getattr(config, name)))
        cfg.ks = int(grid_data.ks)
        cfg.ptop = float(grid_data.ptop)
        cfg.da_min = damping.da_min
        cfg.da_min_c = damping.da_min_c
        grid = _lib.Grid()
        self._keep = []
        for name in _lib.GRID_FIELDS:
            src = {"f0": "fC_agrid", "pfull": "p"}.get(name, name)
            q = None
            for holder in (grid_data, damping):
                try:
                    q = getattr(holder, src)
                    break
                except AttributeError:
                    continue
            if q is None:
                setattr(grid, name, None)
                continue
            t = q.data if hasattr(q, "data") and not isinstance(q, torch.Tensor) else q
            self._keep.append(t)
            setattr(grid, name, t.data_ptr())
        self._a2b_w = torch.as_tensor(a2b_corner_weights(grid_data, g)).to(self.device)
        grid.a2b_w = self._a2b_w.data_ptr()
        n_scratch = self.lib.fv3_scratch_fields()
        self.scratch = torch.zeros(n_scratch * g.ss * g.n_sub, dtype=torch.float64, device=self.device)
        self.c_geom = comm.c_geom
        self.ctx = self.lib.fv3_create(ctypes.byref(self.c_geom), ctypes.byref(cfg), ctypes.byref(grid),
                                       self.scratch.data_ptr(), self.scratch.numel() * 8)
        if not self.ctx:
            raise RuntimeError("fv3_create failed: " + self.lib.fv3_last_error().decode())
        self._cfg, self._grid = cfg, grid

    def stream(self) -> int:
        return self.comm.stream_ptr()

    def call(self, name, *args):
        prof = _lib.PROFILE
        e0 = prof.begin() if prof is not None else None
        rc = getattr(self.lib, name)(self.ctx, *args, self.stream())
        _lib.check(self.lib, rc, name)
        if prof is not None:
            prof.end(name, e0)

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.fv3_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass
