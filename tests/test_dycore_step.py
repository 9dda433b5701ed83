"""Full DynamicalCore.step_dynamics on the c12 baroclinic case, all 6 subdomains in ONE process (batched), compared
with the final state of the unmodified reference (numpy backend).

Runs from the committed subset tests/golden/c12_step (grid + initial state of all 6 ranks, final state of ranks 0 and
3) or, when present, from a full 6-rank dump of oracle/refshim/gen_golden.py (PACE_B200_STEP_CASE selects the case).
The strict multi-substep / layout (2,2) / do_sat_adj cases are in tests/test_step_strict.py.
"""
import os
from datetime import timedelta

import numpy as np
import pytest
import torch

from tests import helpers as H

CASE = os.environ.get("PACE_B200_STEP_CASE", "c12")
# per-field tolerances after ONE step: relative 1e-10 OR below the absolute floor calibrated by the reference itself
# (tests/savepoint/thresholds/fv_dynamics.yaml: round-off sensitivity of one c12 step)
TOL = {"default": (1e-10, 1e-13), "u": (2.1e-9, 1e-11), "v": (2.1e-9, 1e-11), "w": (6.5e-6, 1.5e-12), "ua": (2.1e-9, 1e-11),
       "va": (2.1e-9, 1e-11), "omga": (1e-9, 1e-11), "uc": (1e-8, 1e-11), "vc": (1e-8, 1e-11), "q_con": (1e-9, 1e-15),
       "mfxd": (1e-9, 1e-3), "mfyd": (1e-9, 1e-3), "cxd": (1e-9, 1e-11), "cyd": (1e-9, 1e-11), "diss_estd": (1e-6, 1e-10)}
FIELDS = ["u", "v", "w", "delz", "delp", "pt", "pe", "peln", "pk", "pkz", "ps", "qvapor", "q_con", "omga", "ua", "va"]


def _load(case):
    import json

    base = os.path.join(H.CACHE, case)
    if not os.path.exists(os.path.join(base, "state1_rank5.npz")):
        base = os.path.join(H.GOLDEN, case + "_step")  # committed subset: state1 of ranks 0 and 3 only
        if not os.path.exists(os.path.join(base, "state1_rank0.npz")):
            pytest.skip("reference dump not available")
    meta = json.load(open(os.path.join(base, "meta.json")))
    if meta["layout"] != 1:
        pytest.skip("layout 1 only")
    grids, s0, s1 = [], [], {}
    for r in range(6):
        grids.append(dict(np.load(os.path.join(base, f"grid_rank{r}.npz"))))
        s0.append(dict(np.load(os.path.join(base, f"state0_rank{r}.npz"))))
        p1 = os.path.join(base, f"state1_rank{r}.npz")
        if os.path.exists(p1):
            s1[r] = dict(np.load(p1))
    return meta, grids, s0, s1


def build_dycore(meta, grids, s0, dev=None):
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.dycore_state import DycoreState
    from pace_b200.fv3core.runtime import Runtime
    from pace_b200.fv3core.stencil_factory import GridIndexing, StencilFactory
    from pace_b200.fv3core.stencils.fv_dynamics import DynamicalCore
    from pace_b200.util.grid.helper import DampingCoefficients, GridData

    nx = meta["nx"]
    comm, qf = H.make_comm(nx, 1, 79, dev)
    gd = GridData.from_arrays(qf, grids)
    damp = DampingCoefficients.from_arrays(qf, grids)
    cfg = baroclinic_config(nx, (1, 1), n_split=meta.get("n_split", 1), k_split=meta.get("k_split", 1))
    rt = Runtime(comm, qf, gd, damp, cfg)
    sf = StencilFactory(None, GridIndexing.from_sizer_and_communicator(qf.sizer, comm), rt)
    state = DycoreState.init_from_numpy_arrays(s0, qf)
    dycore = DynamicalCore(comm, gd, sf, qf, damp, cfg, state.phis, state, timedelta(seconds=cfg.dt_atmos))
    return dycore, state


def _run_step():
    meta, grids, s0, s1 = _load(CASE)
    if meta.get("capture_step", 0) != 0 or meta.get("nsteps", 1) != 1:
        pytest.skip("needs a one-step dump")
    dycore, state = build_dycore(meta, grids, s0)
    dycore.step_dynamics(state)
    H.sync()
    nx = meta["nx"]
    c = slice(3, 3 + nx)
    out = state.as_numpy()
    failures = []
    for name in FIELDS:
        rel, floor = TOL.get(name, TOL["default"])
        for r in sorted(s1):
            a, b = out[name][r], s1[r][name]
            if a.ndim == 3:
                nk = 80 if name in ("pe", "peln", "pk") else 79
                ii = slice(3, 3 + nx + (1 if name == "v" else 0))
                jj = slice(3, 3 + nx + (1 if name == "u" else 0))
                a, b = a[ii, jj, :nk], b[ii, jj, :nk]
            else:
                a, b = a[c, c], b[c, c]
            m = H.ref_metric(a, b)
            bad = (m > rel) & (np.abs(a - b) > floor)
            if bad.any():
                failures.append(f"{name} rank {r}: {int(bad.sum())} pts, worst rel {m[bad].max():.2e}, worst abs {np.abs(a - b)[bad].max():.2e}")
    assert not failures, "\n".join(failures)
    # total dry mass is conserved to round-off
    area = np.stack([g["area"][c, c] for g in grids])
    m0 = sum((s0[r]["delp"][c, c, :79] * area[r][:, :, None]).sum() for r in range(6))
    m1 = sum((out["delp"][r][c, c, :79] * area[r][:, :, None]).sum() for r in range(6))
    assert abs(m1 - m0) / m0 < 1e-13


def test_step_dynamics_matches_reference_c12_hostsim():
    if torch.cuda.is_available():
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _run_step()


@pytest.mark.gpu
def test_step_dynamics_matches_reference_c12_gpu():
    _run_step()
