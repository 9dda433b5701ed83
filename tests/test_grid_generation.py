"""pace_b200.util.grid.generation and fv3core.initialization.baroclinic against the reference's own arrays.

Golden data: metric terms and initial state dumped from the unmodified reference (MetricTerms / init_baroclinic_state,
numpy) by oracle/refshim/gen_golden.py (c12, layout 1: tests/golden/c12_step) and oracle/refshim/gen_grid.py
(c24, layout 2, ranks 0/7/13/22: tests/golden/grid_c24L2).
The generator works per cube tile and cuts subdomains out of it, so cells a reference rank cannot compute (outermost
padding row/column, and halo rows of interior subdomain edges, where the reference holds 0 / +-1e8 / NaN placeholders)
hold the neighbour's real value here; those placeholder cells are excluded from the comparison.
"""
import os

import numpy as np
import pytest

from pace_b200.fv3core.initialization import baroclinic
from pace_b200.util.grid import generation
from tests import helpers as H

PLACEHOLDERS = (0.0, 1.0e8, -1.0e8, 1.0e-8, -1.0e-8)


def _compare(mine, ref, name, rel=1e-9, floor=1e-13):
    a, b = np.asarray(mine, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    with np.errstate(invalid="ignore", divide="ignore"):
        placeholder = np.isnan(b) | np.isinf(b) | np.isin(b, PLACEHOLDERS)
        ok = placeholder | (np.abs(a - b) <= rel * np.abs(b) + floor * max(1.0, float(np.nanmax(np.abs(np.where(placeholder, 0, b))))))
    assert ok.all(), f"{name}: {int((~ok).sum())} points differ, first {np.argwhere(~ok)[0]}"
    return int((~placeholder).sum())


def _grid_case(case_dir, N, layout, ranks):
    mine = generation.generate(N, layout, ranks)
    for g, r in zip(mine, ranks):
        ref = np.load(os.path.join(H.GOLDEN, case_dir, f"grid_rank{r}.npz"))
        for k in ref.files:
            if k in ("ks",):
                continue
            assert k in g, f"generator lacks {k}"
            n_real = _compare(g[k], ref[k], f"{case_dir} rank {r} {k}")
            if ref[k].ndim >= 2 and k not in ("edge_w", "edge_e"):
                n = N // layout
                assert n_real >= (n - 1) * (n - 1), f"{k}: too few comparable points ({n_real})"


def test_metric_terms_match_reference_c12():
    _grid_case("c12_step", 12, 1, list(range(6)))


def test_metric_terms_match_reference_c24_layout2():
    _grid_case("grid_c24L2", 24, 2, [0, 7, 13, 22])


def test_edge_factors_only_on_tile_edges():
    g = generation.generate(24, 2, [0, 3])
    # rank 0 = south-west subdomain: west and south factors set, east and north are the 1e8 placeholder
    assert (g[0]["edge_w"][0, 4:15] < 1).all() and (g[0]["edge_s"][4:15] < 1).all()
    assert (g[0]["edge_e"][:, 3:16] == 1e8).all() and (g[0]["edge_n"][3:16] == 1e8).all()
    assert (g[1]["edge_e"][0, 3:15] < 1).all() and (g[1]["edge_w"][:, 3:16] == 1e8).all()


def test_total_area_is_the_sphere():
    for N in (12, 48):
        t = generation.generate_tiles(N)
        area = t["area"][:, 3:3 + N, 3:3 + N].sum()
        assert abs(area / (4 * np.pi * generation.RADIUS ** 2) - 1) < 1e-12


def _state_compare(arrays, s, ref, n, names, interior_only):
    for k in names:
        a, b = arrays[k][s], ref[k]
        if interior_only:
            sl = (slice(3, 3 + n + (1 if k == "v" else 0)), slice(3, 3 + n + (1 if k == "u" else 0)))
            a, b = a[sl], b[sl]
        floor = 1e-12 if k in ("u", "v") else 1e-13 * max(1.0, float(np.abs(b).max()))
        bad = np.abs(a - b) > 1e-12 * np.abs(b) + floor
        assert not bad.any(), f"{k}: {int(bad.sum())} points differ, first {np.argwhere(bad)[0]}"


def test_baroclinic_state_matches_reference_c24_layout2():
    ranks = [7, 22]
    grids = generation.generate(24, 2, ranks)
    st = baroclinic.baroclinic_arrays(grids)
    for s, r in enumerate(ranks):
        ref = np.load(os.path.join(H.GOLDEN, "grid_c24L2", f"state0_rank{r}.npz"))
        _state_compare(st, s, ref, 12, ref.files, interior_only=True)


def test_init_baroclinic_state_with_halos_matches_reference_c12(device):
    """Full path: generated grid -> GridData -> init_baroclinic_state (device upload + halo exchange of phis, u, v)."""
    from pace_b200.util.grid.helper import GridData

    comm, qf = H.make_comm(12, 1, 79, device)
    gd = GridData.new_from_generation(qf, comm)
    state = baroclinic.init_baroclinic_state(gd, qf, adiabatic=False, hydrostatic=False, moist_phys=True, comm=comm)
    H.sync()
    out = state.as_numpy()
    for r in range(6):
        ref = np.load(os.path.join(H.GOLDEN, "c12_step", f"state0_rank{r}.npz"))
        _state_compare(out, r, ref, 12, [k for k in ref.files if k in out], interior_only=False)


@pytest.mark.gpu
def test_init_baroclinic_state_gpu(device):
    test_init_baroclinic_state_with_halos_matches_reference_c12(device)
