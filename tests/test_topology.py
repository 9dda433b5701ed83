"""Known-answer tests of the cubed-sphere topology against the REFERENCE (fixtures under tests/golden/topology):

 * `partitioner_boundaries.json::hand_recorded` — the hand-recorded (from_rank, to_rank, n_clockwise_rotations) tables of
   /root/reference/util/tests/test_partitioner_boundaries.py:34-735 (1x1, 2x2 and the 3x3 "difficult cases"), extracted
   by tests/golden/make_partitioner_tables.py;
 * `partitioner_boundaries.json::reference_partitioner` — every (boundary_type, rank) of the reference's
   CubedSpherePartitioner.boundary for layouts 1-4 (same script, reference imported through oracle/refshim);
 * `halo_known_answers.npz` — the arrays the reference's own halo updater produces from index-encoded fields
   (tests/golden/make_halo_known_answers.py), which pin source rank, source point, component swap and sign of every
   halo point for every staggering used on the hot path.
"""
import json
import os

import numpy as np
import pytest

from pace_b200.util import topology as T
from tests import helpers as H

TOPO = os.path.join(H.GOLDEN, "topology")
BT = {n: getattr(T, n) for n in ("WEST", "EAST", "NORTH", "SOUTH", "NORTHWEST", "NORTHEAST", "SOUTHWEST", "SOUTHEAST")}
STAG = {"x": 1, "y": 1, "x_interface": 0, "y_interface": 0}


def _tables():
    return json.load(open(os.path.join(TOPO, "partitioner_boundaries.json")))


@pytest.mark.parametrize("which", ["hand_recorded", "reference_partitioner"])
def test_neighbour_matches_reference_tables(which):
    rows = _tables()[which]
    assert len(rows) >= 244
    bad = []
    for r in rows:
        dec = T.Decomposition(4, r["layout"])
        got = dec.neighbour(BT[r["boundary"]], r["from_rank"])
        if r["to_rank"] is None:
            if got is not None:
                bad.append((r, got))
            continue
        if got is None or got[0] != r["to_rank"] or (r["rotations"] is not None and got[1] % 4 != r["rotations"] % 4):
            bad.append((r, got))
    assert not bad, f"{len(bad)} of {len(rows)} rows differ, first: {bad[:3]}"


def _encode(shape, rank, code):
    i, j = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    return code + rank * 1e4 + i * 100 + j


def _cases():
    z = np.load(os.path.join(TOPO, "halo_known_answers.npz"))
    names = sorted({k.rsplit(".", 1)[0] for k in z.files})
    return z, names


CASE_DIMS = {
    "scalar_cell_h3": (("x", "y"), None, 3, "halo"), "scalar_cell_h2": (("x", "y"), None, 2, "halo"),
    "scalar_cell_h1": (("x", "y"), None, 1, "halo"), "scalar_corner_h3": (("x_interface", "y_interface"), None, 3, "halo"),
    "scalar_zi_h3": (("x", "y"), None, 3, "halo"),
    "vector_dgrid_h3": (("x", "y_interface"), ("x_interface", "y"), 3, "halo"),
    "vector_cgrid_h3": (("x_interface", "y"), ("x", "y_interface"), 3, "halo"),
    "vector_agrid_h3": (("x", "y"), ("x", "y"), 3, "halo"),
    "vector_dgrid_h1": (("x", "y_interface"), ("x_interface", "y"), 1, "halo"),
    "sync_interfaces_dgrid": (("x", "y_interface"), ("x_interface", "y"), 0, "interface"),
}


@pytest.mark.parametrize("layout", [1, 2, 3])
@pytest.mark.parametrize("case", sorted(CASE_DIMS))
def test_halo_table_matches_reference_halo_updater(layout, case):
    z, names = _cases()
    key = f"L{layout}.{case}"
    assert key in names
    dx, dy, nh, mode = CASE_DIMS[case]
    ref_x = z[key + ".x"]
    ref_y = z[key + ".y"] if dy is not None else None
    nranks = ref_x.shape[0]
    dec = T.Decomposition(4, layout)
    assert dec.total_ranks == nranks
    fx = [_encode(ref_x.shape[1:], r, 1.0) for r in range(nranks)]
    fy = [_encode(ref_y.shape[1:], r, 0.5) for r in range(nranks)] if dy is not None else None
    sx = (STAG[dx[0]], STAG[dx[1]])
    sy = (STAG[dy[0]], STAG[dy[1]]) if dy is not None else None
    table = T.build_halo_table(dec, nh, sx, sy, mode=mode)
    T.apply_table_numpy(table, fx, fy)
    np.testing.assert_array_equal(np.stack(fx), ref_x)
    if dy is not None:
        np.testing.assert_array_equal(np.stack(fy), ref_y)
