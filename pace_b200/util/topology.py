"""Cubed-sphere topology and halo-exchange index tables (host-side, numpy).

Replaces, for the halo path, the reference's `CubedSpherePartitioner.boundary` case analysis
(util/pace/util/partitioner.py:406-606), `Boundary.send_slice/recv_slice` (boundary.py:39-113,
_boundary_utils.py:58-101) and the rotation helpers (rotate.py:4-50) by ONE geometric rule:

  every grid point lives at doubled tile coordinates P = (2*gi + ox, 2*gj + oy), ox/oy = 1 for a cell-centred
  dimension and 0 for an interface dimension, so a tile spans [0, 2N] in both directions.  A halo point with
  exactly one coordinate outside [0, 2N] is carried into the neighbouring tile by the integer affine map
  P' = A.P + b of that tile edge (EDGE_MAPS below); vector components transform with A^T.  Points outside in
  both directions lie beyond a cube corner and have no source (the reference leaves them untouched too).

The result is a flat gather table (dst subdomain/offset <- src subdomain/offset, sign, component) that one
CUDA kernel consumes for all fields, all levels and all local subdomains of an exchange.
Known-answer checks: tests/test_topology.py replays the hand-recorded tables of the reference
(util/tests/test_partitioner_boundaries.py:34-735).
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

WEST, EAST, NORTH, SOUTH, NORTHWEST, NORTHEAST, SOUTHWEST, SOUTHEAST = range(8)
_DIRS = {
    WEST: (-1, 0), EAST: (1, 0), NORTH: (0, 1), SOUTH: (0, -1),
    NORTHWEST: (-1, 1), NORTHEAST: (1, 1), SOUTHWEST: (-1, -1), SOUTHEAST: (1, -1),
}

_I = ((1, 0), (0, 1))
_RP = ((0, 1), (-1, 0))   # (X, Y) -> ( Y, -X)
_RM = ((0, -1), (1, 0))   # (X, Y) -> (-Y,  X)
# EDGE_MAPS[tile parity][edge] = (tile delta, A, b in units of M = 2N, n_clockwise_rotations of the reference)
EDGE_MAPS = {
    0: {EAST: (1, _I, (-1, 0), 0), NORTH: (2, _RP, (-1, 1), 3), WEST: (-2, _RM, (1, 1), 1), SOUTH: (-1, _I, (0, 1), 0)},
    1: {NORTH: (1, _I, (0, -1), 0), EAST: (2, _RM, (1, -1), 1), SOUTH: (-2, _RP, (1, 1), 3), WEST: (-1, _I, (1, 0), 0)},
}


def _cross(tile, X, Y, M):
    """Carry doubled tile coordinates (arrays) across at most one tile edge.

    Returns (tile', X', Y', A, valid) with A the 2x2 map used (per point, as 4 int arrays) and valid=False
    for points beyond a cube corner.
    """
    X = np.asarray(X)
    Y = np.asarray(Y)
    tile = np.broadcast_to(np.asarray(tile), X.shape).copy()
    out_w, out_e, out_s, out_n = X < 0, X > M, Y < 0, Y > M
    nout = out_w.astype(int) + out_e + out_s + out_n
    valid = nout <= 1
    Xo, Yo, To = X.copy(), Y.copy(), tile.copy()
    a00 = np.ones_like(X)
    a01 = np.zeros_like(X)
    a10 = np.zeros_like(X)
    a11 = np.ones_like(X)
    for parity in (0, 1):
        for edge, mask in ((WEST, out_w), (EAST, out_e), (SOUTH, out_s), (NORTH, out_n)):
            m = mask & valid & ((tile % 2) == parity)
            if not m.any():
                continue
            dt, A, b, _ = EDGE_MAPS[parity][edge]
            Xo[m] = A[0][0] * X[m] + A[0][1] * Y[m] + b[0] * M
            Yo[m] = A[1][0] * X[m] + A[1][1] * Y[m] + b[1] * M
            To[m] = (tile[m] + dt) % 6
            a00[m], a01[m], a10[m], a11[m] = A[0][0], A[0][1], A[1][0], A[1][1]
    return To, Xo, Yo, (a00, a01, a10, a11), valid


@dataclass(frozen=True)
class Decomposition:
    """6 * layout^2 square subdomains of nx*nx cells; rank = tile*L^2 + tj*L + ti (partitioner.py:31-38,746-752)."""

    nx: int          # cells per subdomain side
    layout: int = 1  # subdomains per tile side

    @property
    def total_ranks(self):
        return 6 * self.layout * self.layout

    @property
    def n_tile(self):
        return self.nx * self.layout

    def tile_of(self, rank):
        return rank // (self.layout * self.layout)

    def subtile_index(self, rank):
        w = rank % (self.layout * self.layout)
        return w // self.layout, w % self.layout  # (tj, ti)

    def rank_of(self, tile, tj, ti):
        return tile * self.layout * self.layout + tj * self.layout + ti

    def edge_flags(self, rank):
        """(west, east, south, north): is this subdomain on that edge of its tile."""
        tj, ti = self.subtile_index(rank)
        L = self.layout
        return ti == 0, ti == L - 1, tj == 0, tj == L - 1

    def neighbour(self, boundary_type: int, rank: int) -> Optional[Tuple[int, int]]:
        """(to_rank, n_clockwise_rotations) of the reference's SimpleBoundary, or None at a cube corner."""
        dx, dy = _DIRS[boundary_type]
        tj, ti = self.subtile_index(rank)
        n2 = 2 * self.nx
        M = 2 * self.n_tile
        cx = np.array([(ti + dx) * n2 + self.nx])
        cy = np.array([(tj + dy) * n2 + self.nx])
        t, X, Y, A, valid = _cross(self.tile_of(rank), cx, cy, M)
        if not valid[0]:
            return None
        to_rank = self.rank_of(int(t[0]), int(Y[0]) // n2, int(X[0]) // n2)
        a = (int(A[0][0]), int(A[1][0]), int(A[2][0]), int(A[3][0]))
        rot = {(1, 0, 0, 1): 0, _RM[0] + _RM[1]: 1, _RP[0] + _RP[1]: 3}[a]
        if boundary_type >= NORTHWEST and rot == 0:
            # corners reached through two rotated edges report the summed rotation in the reference
            # (partitioner.py:_get_corner); a single crossing already gives the right data mapping.
            pass
        return to_rank, rot


@dataclass
class HaloTable:
    """Flat gather table of one exchange pattern; all arrays have one entry per destination halo point."""

    dst_rank: np.ndarray   # int32
    dst_comp: np.ndarray   # int8   0 = x-field, 1 = y-field of a vector pair (0 for scalars)
    dst_i: np.ndarray      # int32  local storage index i (incl. halo origin)
    dst_j: np.ndarray
    src_rank: np.ndarray
    src_comp: np.ndarray
    src_i: np.ndarray
    src_j: np.ndarray
    sign: np.ndarray       # float64 +-1

    def __len__(self):
        return len(self.dst_rank)

    def select(self, mask):
        return HaloTable(*[getattr(self, f)[mask] for f in self.__dataclass_fields__])


def _points(decomp, rank, ox, oy, n_halo, halo, mode):
    """Local storage indices (i, j) and doubled tile coordinates of the destination points of `rank`."""
    nx = decomp.nx
    n2 = 2 * nx
    tj, ti = decomp.subtile_index(rank)
    npx = nx + (1 - ox)
    npy = nx + (1 - oy)
    if mode == "halo":
        i = np.arange(-n_halo, npx + n_halo)
        j = np.arange(-n_halo, npy + n_halo)
        I, J = np.meshgrid(i, j, indexing="ij")
        inside = (I >= 0) & (I < npx) & (J >= 0) & (J < npy)
        I, J = I[~inside], J[~inside]
    elif mode == "interface":
        # shared-edge synchronisation (halo_updater.py:385-416): the north row of a y-interface field and the
        # east column of an x-interface field are overwritten with the neighbour's south row / west column.
        if ox == 1 and oy == 0:
            I = np.arange(0, npx)
            J = np.full_like(I, npy - 1)
        elif ox == 0 and oy == 1:
            J = np.arange(0, npy)
            I = np.full_like(J, npx - 1)
        else:
            raise ValueError("interface synchronisation needs exactly one interface dimension")
    else:
        raise ValueError(mode)
    Xl = 2 * I + ox
    Yl = 2 * J + oy
    return I + halo, J + halo, Xl, Yl, Xl + ti * n2, Yl + tj * n2


def build_halo_table(decomp: Decomposition, n_halo: int, stagger_x=(1, 1), stagger_y=None, halo: int = 3,
                     mode: str = "halo") -> HaloTable:
    """Gather table for a scalar field (stagger_y None) or a vector pair.

    stagger = (ox, oy): 1 = cell-centred, 0 = interface in that direction.  For a pair, `stagger_x` is the
    staggering of the x-component field (e.g. D-grid u: (1, 0); C-grid uc: (0, 1)) and `stagger_y` that of the
    y-component field.
    """
    nx = decomp.nx
    n2 = 2 * nx
    M = 2 * decomp.n_tile
    cols: Dict[str, List[np.ndarray]] = {f: [] for f in HaloTable.__dataclass_fields__}
    comps = [(0, stagger_x)] + ([(1, stagger_y)] if stagger_y is not None else [])
    for rank in range(decomp.total_ranks):
        tile = decomp.tile_of(rank)
        tj, ti = decomp.subtile_index(rank)
        for comp, (ox, oy) in comps:
            si, sj, Xl, Yl, Xg, Yg = _points(decomp, rank, ox, oy, n_halo, halo, mode)
            # which neighbour (relative subtile offset) serves each point: the reference takes edge strips over
            # the full (interface-extended) compute extent and corners beyond it.
            if mode == "halo":
                dx = np.where(Xl < 0, -1, np.where(Xl > n2, 1, 0))
                dy = np.where(Yl < 0, -1, np.where(Yl > n2, 1, 0))
            else:
                dx = np.where(np.full(Xl.shape, ox == 0), 1, 0)
                dy = np.where(np.full(Xl.shape, oy == 0), 1, 0)
            # representative interior point of the serving subdomain -> its tile / rank
            cx = (ti + dx) * n2 + nx
            cy = (tj + dy) * n2 + nx
            t2, CX, CY, A, valid = _cross(tile, cx, cy, M)
            # the destination point itself, moved with the SAME edge map as the serving subdomain's centre
            a00, a01, a10, a11 = A
            out_w, out_e, out_s, out_n = cx < 0, cx > M, cy < 0, cy > M
            PX, PY = Xg.copy(), Yg.copy()
            for parity in (0, 1):
                if tile % 2 != parity:
                    continue
                for edge, mask in ((WEST, out_w), (EAST, out_e), (SOUTH, out_s), (NORTH, out_n)):
                    m = mask & valid
                    if not m.any():
                        continue
                    _, Am, b, _ = EDGE_MAPS[parity][edge]
                    PX[m] = Am[0][0] * Xg[m] + Am[0][1] * Yg[m] + b[0] * M
                    PY[m] = Am[1][0] * Xg[m] + Am[1][1] * Yg[m] + b[1] * M
            src_ti = CX // n2
            src_tj = CY // n2
            src_rank = t2 * decomp.layout * decomp.layout + src_tj * decomp.layout + src_ti
            lx = PX - src_ti * n2
            ly = PY - src_tj * n2
            # my components = A^T . their components; the row of A^T for my component `comp`
            if stagger_y is None:
                src_comp = np.zeros_like(lx)
                sign = np.ones(lx.shape)
            else:
                if comp == 0:
                    cx_, cy_ = a00, a10   # ux = a00*ux' + a10*uy'
                else:
                    cx_, cy_ = a01, a11   # uy = a01*ux' + a11*uy'
                src_comp = np.where(cx_ != 0, 0, 1)
                sign = np.where(cx_ != 0, cx_, cy_).astype(np.float64)
            v = valid
            cols["dst_rank"].append(np.full(v.sum(), rank))
            cols["dst_comp"].append(np.full(v.sum(), comp))
            cols["dst_i"].append(si[v])
            cols["dst_j"].append(sj[v])
            cols["src_rank"].append(src_rank[v])
            cols["src_comp"].append(src_comp[v])
            cols["src_i"].append(lx[v] // 2 + halo)
            cols["src_j"].append(ly[v] // 2 + halo)
            cols["sign"].append(sign[v])
    out = {k: np.concatenate(v) for k, v in cols.items()}
    return HaloTable(
        out["dst_rank"].astype(np.int32), out["dst_comp"].astype(np.int8), out["dst_i"].astype(np.int32),
        out["dst_j"].astype(np.int32), out["src_rank"].astype(np.int32), out["src_comp"].astype(np.int8),
        out["src_i"].astype(np.int32), out["src_j"].astype(np.int32), out["sign"].astype(np.float64),
    )


def apply_table_numpy(table: HaloTable, fields_x, fields_y=None):
    """Apply a gather table on host arrays (used for grid generation and by the CPU tests, never on the hot path).

    fields_x / fields_y: lists indexed by rank of arrays [i, j, ...].  All sources are read before any write.
    """
    comps = [fields_x, fields_y]
    vals = np.empty((len(table),) + tuple(fields_x[0].shape[2:]), dtype=fields_x[0].dtype)
    for r in np.unique(table.src_rank):
        for c in (0, 1):
            m = (table.src_rank == r) & (table.src_comp == c)
            if m.any():
                vals[m] = comps[c][r][table.src_i[m], table.src_j[m]]
    vals = vals * table.sign.reshape((-1,) + (1,) * (vals.ndim - 1)).astype(vals.dtype)
    for r in np.unique(table.dst_rank):
        for c in (0, 1):
            m = (table.dst_rank == r) & (table.dst_comp == c)
            if m.any():
                comps[c][r][table.dst_i[m], table.dst_j[m]] = vals[m]
