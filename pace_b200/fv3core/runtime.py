"""Native context shared by all stage objects of one process: geometry, config, metric-term pointers, scratch."""
import ctypes

import torch

from .. import _lib
from ..util.grid.helper import DampingCoefficients, GridData


class Runtime:
    def __init__(self, comm, quantity_factory, grid_data: GridData, damping: DampingCoefficients, config):
        self.lib = _lib.load()
        self.comm = comm
        self.qf = quantity_factory
        self.grid_data = grid_data
        self.damping = damping
        self.config = config
        self.device = quantity_factory.device
        if self.device.type != "cuda" and not _lib.hostsim_requested():
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
        rc = getattr(self.lib, name)(self.ctx, *args, self.stream())
        _lib.check(self.lib, rc, name)

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.fv3_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass
