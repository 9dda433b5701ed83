"""Cubed-sphere communicator and halo updater for one-process-per-GPU, many-subdomains-per-process runs.

Keeps the reference API for the halo path (util/pace/util/communicator.py:331-555, halo_updater.py:29-303):
    comm.get_scalar_halo_updater(specs) / get_vector_halo_updater(specs_x, specs_y)
    updater.start(quantities_x, quantities_y=None); updater.wait(); updater.update(...)
    comm.halo_update / vector_halo_update / synchronize_vector_interfaces
Design: the gather table of pace_b200.util.topology is split by where source and destination live:
  same process  -> ONE `fv3_halo_gather` launch that reads the neighbour subdomain's array directly (rotation,
                   component swap and sign folded into the table; no staging buffer);
  other process -> ONE `fv3_halo_pack` launch per exchange into a buffer holding one contiguous segment per
                   peer, NCCL send/recv of the segments (torch.distributed batch_isend_irecv), ONE
                   `fv3_halo_unpack` launch on wait().
"""
import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .. import constants as c
from . import topology
from .quantity import Quantity, QuantityHaloSpec
from .sizer import Geometry


class TilePartitioner:
    def __init__(self, layout: Tuple[int, int]):
        if layout[0] != layout[1]:
            raise NotImplementedError("only square layouts are supported (as partitioner.py:_ensure_square_layout)")
        self.layout = tuple(layout)

    @property
    def total_ranks(self):
        return self.layout[0] * self.layout[1]


class CubedSpherePartitioner:
    def __init__(self, tile: TilePartitioner):
        self.tile = tile
        self.layout = tile.layout

    @property
    def total_ranks(self):
        return 6 * self.tile.total_ranks

    def tile_index(self, rank):
        return rank // self.tile.total_ranks

    def boundary(self, boundary_type: int, rank: int, nx: int = 4):
        """(to_rank, n_clockwise_rotations) or None — see topology.Decomposition.neighbour."""
        return topology.Decomposition(nx, self.layout[0]).neighbour(boundary_type, rank)


class ProcessComm:
    """Process-level communication: which GPU process owns which subdomain ranks, and p2p between processes.

    size == 1: everything is device-local.  size > 1: torch.distributed (NCCL on GPUs, gloo in CPU tests).
    """

    def __init__(self, rank: int = 0, size: int = 1, group=None):
        self.rank = rank
        self.size = size
        self.group = group

    @classmethod
    def from_torch_distributed(cls, group=None):
        import torch.distributed as dist

        return cls(dist.get_rank(group), dist.get_world_size(group), group)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size


def owner_of(rank: int, total_ranks: int, n_proc: int) -> int:
    """Process owning subdomain `rank`: contiguous blocks of total_ranks / n_proc."""
    per = total_ranks // n_proc
    return rank // per


class HaloUpdater:
    """start/wait state machine of halo_updater.py:217-303 over precomputed device gather tables."""

    def __init__(self, comm: "CubedSphereCommunicator", table: topology.HaloTable, vector: bool, two_d: bool, nlev: int):
        self._comm = comm
        self._vector = vector
        self._nlev = nlev
        self._inflight = False
        g = comm.geometry
        dev = comm.device
        pc = comm.process_comm
        total = comm.partitioner.total_ranks
        first = comm.first_rank
        ss = g.ss2 if two_d else g.ss
        dst_proc = table.dst_rank // comm.ranks_per_process
        src_proc = table.src_rank // comm.ranks_per_process
        me = pc.rank

        def off(rank, i, j):
            return ((rank - first).astype(np.int64) * ss + j.astype(np.int64) * g.sj + i.astype(np.int64))

        def dev_t(a, dtype):
            return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(dev)

        # device-local part
        m = (dst_proc == me) & (src_proc == me)
        self._n_local = int(m.sum())
        self._loc = dict(
            dst_off=dev_t(off(table.dst_rank[m], table.dst_i[m], table.dst_j[m]), torch.int64),
            src_off=dev_t(off(table.src_rank[m], table.src_i[m], table.src_j[m]), torch.int64),
            dst_comp=dev_t(table.dst_comp[m], torch.int8),
            src_comp=dev_t(table.src_comp[m], torch.int8),
            sign=dev_t(table.sign[m], torch.float64),
        )
        # remote parts: canonical order = table order restricted to the (src_proc, dst_proc) pair
        self._send = []  # (peer, n, src_off, src_comp, sign)
        self._recv = []  # (peer, n, dst_off, dst_comp)
        for peer in range(pc.size):
            if peer == me:
                continue
            ms = (src_proc == me) & (dst_proc == peer)
            if ms.any():
                self._send.append((peer, int(ms.sum()),
                                   dev_t(off(table.src_rank[ms], table.src_i[ms], table.src_j[ms]), torch.int64),
                                   dev_t(table.src_comp[ms], torch.int8), dev_t(table.sign[ms], torch.float64)))
            mr = (dst_proc == me) & (src_proc == peer)
            if mr.any():
                self._recv.append((peer, int(mr.sum()),
                                   dev_t(off(table.dst_rank[mr], table.dst_i[mr], table.dst_j[mr]), torch.int64),
                                   dev_t(table.dst_comp[mr], torch.int8)))
        self._bufs: Dict[int, tuple] = {}
        self._ptr_cache: Dict[tuple, torch.Tensor] = {}
        self._pending = None
        _ = total

    # ------------------------------------------------------------------
    def _field_ptrs(self, qx: Sequence[Quantity], qy: Optional[Sequence[Quantity]]):
        key = tuple(q.ptr for q in qx) + (tuple(q.ptr for q in qy) if qy else ())
        t = self._ptr_cache.get(key)
        if t is None:
            t = torch.tensor(list(key), dtype=torch.int64).to(self._comm.device)
            self._ptr_cache[key] = t
        return t

    def _buffers(self, n_fields):
        b = self._bufs.get(n_fields)
        if b is None:
            dev = self._comm.device
            sb = [torch.empty(n_fields * self._nlev * n, dtype=torch.float64, device=dev) for (_, n, *_r) in self._send]
            rb = [torch.empty(n_fields * self._nlev * n, dtype=torch.float64, device=dev) for (_, n, *_r) in self._recv]
            b = (sb, rb)
            self._bufs[n_fields] = b
        return b

    def start(self, quantities_x: Sequence[Quantity], quantities_y: Optional[Sequence[Quantity]] = None):
        if self._inflight:
            raise RuntimeError("HaloUpdater.start called twice without a wait() in between")
        if self._vector != (quantities_y is not None):
            raise ValueError("vector updater needs quantities_y, scalar updater must not get it")
        lib = _lib.load()
        comm = self._comm
        n_fields = len(quantities_x)
        if self._vector and len(quantities_y) != n_fields:
            raise ValueError("quantities_x and quantities_y must have the same length")
        ptrs = self._field_ptrs(quantities_x, quantities_y)
        stream = comm.stream_ptr()
        gp = ctypes.byref(comm.c_geom)
        self._inflight = True
        reqs = []
        prof = _lib.PROFILE
        e0 = prof.begin() if prof is not None else None
        if self._send or self._recv:
            import torch.distributed as dist

            sb, rb = self._buffers(n_fields)
            for (peer, n, src_off, src_comp, sign), buf in zip(self._send, sb):
                _lib.check(lib, lib.fv3_halo_pack(gp, ptrs.data_ptr(), n_fields, self._nlev, src_off.data_ptr(),
                                                  src_comp.data_ptr(), sign.data_ptr(), n, buf.data_ptr(), stream),
                           "fv3_halo_pack")
            ops = [dist.P2POp(dist.isend, buf, peer, comm.process_comm.group) for (peer, *_), buf in zip(self._send, sb)]
            ops += [dist.P2POp(dist.irecv, buf, peer, comm.process_comm.group) for (peer, *_), buf in zip(self._recv, rb)]
            reqs = dist.batch_isend_irecv(ops)
        if self._n_local:
            L = self._loc
            _lib.check(lib, lib.fv3_halo_gather(gp, ptrs.data_ptr(), n_fields, self._nlev, L["dst_off"].data_ptr(),
                                                L["src_off"].data_ptr(), L["dst_comp"].data_ptr(),
                                                L["src_comp"].data_ptr(), L["sign"].data_ptr(), self._n_local, stream),
                       "fv3_halo_gather")
        self._pending = (reqs, ptrs, n_fields)
        if prof is not None:
            prof.end("halo_start", e0)

    def wait(self):
        if not self._inflight:
            raise RuntimeError("HaloUpdater.wait called before start")
        reqs, ptrs, n_fields = self._pending
        if reqs:
            lib = _lib.load()
            for r in reqs:
                r.wait()
            _, rb = self._buffers(n_fields)
            gp = ctypes.byref(self._comm.c_geom)
            stream = self._comm.stream_ptr()
            for (peer, n, dst_off, dst_comp), buf in zip(self._recv, rb):
                _lib.check(lib, lib.fv3_halo_unpack(gp, ptrs.data_ptr(), n_fields, self._nlev, dst_off.data_ptr(),
                                                    dst_comp.data_ptr(), n, buf.data_ptr(), stream), "fv3_halo_unpack")
        self._pending = None
        self._inflight = False

    def update(self, quantities_x, quantities_y=None):
        self.start(quantities_x, quantities_y)
        self.wait()

    def __del__(self):
        if getattr(self, "_inflight", False):
            import warnings

            warnings.warn("HaloUpdater garbage-collected while an exchange was in flight")


def _stagger(dims):
    return (0 if dims[0] == c.X_INTERFACE_DIM else 1, 0 if dims[1] == c.Y_INTERFACE_DIM else 1)


class CubedSphereCommunicator:
    """One per process (GPU).  `rank` is the process rank; the subdomain ranks it owns are `local_ranks`."""

    def __init__(self, comm: ProcessComm, partitioner: CubedSpherePartitioner, nx_tile: int, nz: int,
                 n_halo: int = c.N_HALO_DEFAULT, device="cuda", timer=None):
        self.process_comm = comm
        self.partitioner = partitioner
        total = partitioner.total_ranks
        if total % comm.size:
            raise ValueError(f"{total} subdomains cannot be spread evenly over {comm.size} processes")
        self.ranks_per_process = total // comm.size
        self.first_rank = comm.rank * self.ranks_per_process
        self.local_ranks = list(range(self.first_rank, self.first_rank + self.ranks_per_process))
        L = partitioner.layout[0]
        self.decomposition = topology.Decomposition(nx_tile // L, L)
        edges = []
        for r in self.local_ranks:
            w, e, s, n = self.decomposition.edge_flags(r)
            edges.append(1 * w + 2 * e + 4 * s + 8 * n)
        self.geometry = Geometry(len(self.local_ranks), nx_tile // L, nx_tile // L, nz, n_halo, tuple(edges))
        self.c_geom = self.geometry.to_c()
        self.device = torch.device(device)
        self._tables: Dict[tuple, topology.HaloTable] = {}
        self.timer = timer

    @property
    def rank(self):
        return self.process_comm.rank

    @classmethod
    def from_layout(cls, comm, layout, nx_tile, nz, **kw):
        return cls(comm, CubedSpherePartitioner(TilePartitioner(tuple(layout))), nx_tile, nz, **kw)

    def stream_ptr(self) -> int:
        if self.device.type == "cuda":
            return torch.cuda.current_stream(self.device).cuda_stream
        return 0

    def _table(self, n_halo, sx, sy, mode):
        key = (n_halo, sx, sy, mode)
        t = self._tables.get(key)
        if t is None:
            t = topology.build_halo_table(self.decomposition, n_halo, sx, sy, halo=self.geometry.halo, mode=mode)
            self._tables[key] = t
        return t

    def _check_specs(self, specs):
        if len(specs) == 0:
            raise ValueError("need at least one halo specification")
        s0 = specs[0]
        for s in specs:
            if s.n_points == 0:
                raise ValueError("cannot perform a halo update on zero halo points")
            if s.dims != s0.dims or s.n_points != s0.n_points:
                raise NotImplementedError("all fields of one exchange must share dims and n_points")
        return s0

    @staticmethod
    def _nlev(spec, geometry):
        if len(spec.dims) == 2:
            return 1, True
        return (geometry.nz + 1 if spec.dims[2] == c.Z_INTERFACE_DIM else geometry.nz), False

    def get_scalar_halo_updater(self, specifications: Sequence[QuantityHaloSpec]) -> HaloUpdater:
        s0 = self._check_specs(specifications)
        nlev, two_d = self._nlev(s0, self.geometry)
        return HaloUpdater(self, self._table(s0.n_points, _stagger(s0.dims), None, "halo"), False, two_d, nlev)

    def get_vector_halo_updater(self, specifications_x, specifications_y) -> HaloUpdater:
        sx = self._check_specs(specifications_x)
        sy = self._check_specs(specifications_y)
        nlev, two_d = self._nlev(sx, self.geometry)
        return HaloUpdater(self, self._table(sx.n_points, _stagger(sx.dims), _stagger(sy.dims), "halo"), True, two_d, nlev)

    def get_interface_updater(self, spec_x, spec_y) -> HaloUpdater:
        nlev, two_d = self._nlev(spec_x, self.geometry)
        return HaloUpdater(self, self._table(0, _stagger(spec_x.dims), _stagger(spec_y.dims), "interface"), True, two_d, nlev)

    # convenience one-shot forms (communicator.py:331-470)
    def _spec_of(self, q: Quantity, n_points):
        return QuantityHaloSpec(n_points, (), 8, q.shape[1:], q.origin, q.extent, q.dims, np, np.float64)

    def halo_update(self, quantity, n_points: int):
        qs = [quantity] if isinstance(quantity, Quantity) else list(quantity)
        self.get_scalar_halo_updater([self._spec_of(q, n_points) for q in qs]).update(qs)

    def vector_halo_update(self, x_quantity, y_quantity, n_points: int):
        qx = [x_quantity] if isinstance(x_quantity, Quantity) else list(x_quantity)
        qy = [y_quantity] if isinstance(y_quantity, Quantity) else list(y_quantity)
        self.get_vector_halo_updater([self._spec_of(q, n_points) for q in qx],
                                     [self._spec_of(q, n_points) for q in qy]).update(qx, qy)

    def synchronize_vector_interfaces(self, x_quantity: Quantity, y_quantity: Quantity):
        self.get_interface_updater(self._spec_of(x_quantity, 1), self._spec_of(y_quantity, 1)).update([x_quantity], [y_quantity])
