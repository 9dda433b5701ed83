"""Shared test helpers: golden loading, communicator construction, reference-metric comparison."""
import os

import numpy as np
import torch

from pace_b200 import constants as c
from pace_b200.util.communicator import CubedSphereCommunicator, ProcessComm
from pace_b200.util.sizer import QuantityFactory, SubtileGridSizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CACHE = os.environ.get("PACE_B200_GOLDEN_CACHE", "/tmp/pace_b200_golden")  # uncommitted full reference dumps

D3 = (c.X_DIM, c.Y_DIM, c.Z_DIM)
D3I = (c.X_DIM, c.Y_DIM, c.Z_INTERFACE_DIM)
D2 = (c.X_DIM, c.Y_DIM)


def device():
    return "cuda" if torch.cuda.is_available() else "cpu"


def make_comm(nx_tile, layout=1, nz=79, dev=None, process_comm=None):
    dev = dev or device()
    comm = CubedSphereCommunicator.from_layout(process_comm or ProcessComm(), (layout, layout), nx_tile, nz, device=dev)
    sizer = SubtileGridSizer.from_tile_params(nx_tile, nx_tile, nz, 3, {}, (layout, layout))
    qf = QuantityFactory(sizer, comm.geometry, dev)
    return comm, qf


def golden_path(case, *parts):
    """Committed golden file if present, else the local (uncommitted) cache; None if neither exists."""
    for base in (GOLDEN, CACHE):
        p = os.path.join(base, case, *parts)
        if os.path.exists(p):
            return p
    return None


def load_stage(case, rank, stage):
    p = golden_path(case, f"stage_rank{rank}", stage + ".npz")
    if p is None:
        return None
    z = np.load(p)
    d = {k: z[k] for k in z.files}
    for k in list(d):  # outputs identical to their inputs are not stored in the committed subset
        if k.startswith("in.") and "out." + k[3:] not in d:
            d["out." + k[3:]] = d[k]
    return d


def ref_metric(a, b):
    """2|a-b|/(|a|+|b|) of util/pace/util/testing/comparison.py:6-21 (0 where both are 0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.abs(a) + np.abs(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        m = np.where(denom > 0, 2.0 * np.abs(a - b) / denom, 0.0)
    return m


def assert_close(actual, expected, max_error=1e-13, near_zero=1e-18, name=""):
    """Reference success criterion (comparison.py:24-68): relative metric < max_error OR both |.| < near_zero."""
    actual = np.asarray(actual)
    expected = np.asarray(expected)
    assert actual.shape == expected.shape, (name, actual.shape, expected.shape)
    nan_mismatch = np.isnan(actual) != np.isnan(expected)
    assert not nan_mismatch.any(), f"{name}: NaN pattern differs at {int(nan_mismatch.sum())} points"
    ok = (ref_metric(actual, expected) < max_error) | ((np.abs(actual) < near_zero) & (np.abs(expected) < near_zero))
    ok |= np.isnan(expected)
    if not ok.all():
        bad = np.argwhere(~ok)
        i = tuple(bad[0])
        worst = np.nanmax(np.where(ok, 0, ref_metric(actual, expected)))
        raise AssertionError(
            f"{name}: {len(bad)} of {actual.size} points differ (worst rel {worst:.3e}); first at {i}: "
            f"actual={actual[i]!r} expected={expected[i]!r}")


# ---------------------------------------------------------------------------------------------
# c12 golden case: runtime built from the reference's own grid data

_case_cache = {}


def load_case(case="c12", ranks=(0,), dev=None, **config_overrides):
    """(comm, qf, runtime, stencil_factory) for the golden case, holding the given reference ranks as the
    local subdomains (edge flags taken from the real decomposition of those ranks)."""
    import json

    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.runtime import Runtime
    from pace_b200.fv3core.stencil_factory import GridIndexing, StencilFactory
    from pace_b200.util.grid.helper import DampingCoefficients, GridData
    from pace_b200.util.sizer import Geometry
    from pace_b200.util import topology

    key = (case, tuple(ranks), dev, tuple(sorted(config_overrides.items())))
    if key in _case_cache:
        return _case_cache[key]
    meta_p = golden_path(case, "meta.json")
    if meta_p is None:
        return None
    meta = json.load(open(meta_p))
    nx_tile, layout = meta["nx"], meta["layout"]
    dev = dev or device()
    comm, qf = make_comm(nx_tile, layout, 79, dev)
    dec = topology.Decomposition(nx_tile // layout, layout)
    edges = []
    for r in ranks:
        w, e, s, n = dec.edge_flags(r)
        edges.append(1 * w + 2 * e + 4 * s + 8 * n)
    geom = Geometry(len(ranks), nx_tile // layout, nx_tile // layout, 79, 3, tuple(edges))
    comm.geometry = geom
    comm.c_geom = geom.to_c()
    comm.local_ranks = list(ranks)
    qf = QuantityFactory(qf.sizer, geom, dev)
    grids = []
    for r in ranks:
        z = np.load(golden_path(case, f"grid_rank{r}.npz"))
        grids.append({k: z[k] for k in z.files})
    gd = GridData.from_arrays(qf, grids)
    damp = DampingCoefficients.from_arrays(qf, grids)
    cfg = baroclinic_config(nx_tile, (layout, layout), n_split=meta.get("n_split", 1), k_split=meta.get("k_split", 1),
                            **config_overrides)
    rt = Runtime(comm, qf, gd, damp, cfg)
    sf = StencilFactory(None, GridIndexing.from_sizer_and_communicator(qf.sizer, comm), rt)
    out = (comm, qf, rt, sf)
    _case_cache[key] = out
    return out


def to_q(qf, arrays, dims=None):
    """Quantity from per-subdomain reference arrays ([i, j(, k)] each); dims guessed from ndim if omitted."""
    arrs = [np.asarray(a, dtype=np.float64) for a in arrays]
    if dims is None:
        dims = D3 if arrs[0].ndim == 3 else D2
    return qf.from_array(np.stack(arrs), dims, "")


def sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()
