"""STRICT full-step parity cases: the REFERENCE's own grid and initial state go in (bit for bit), ONE
`DynamicalCore.step_dynamics` runs on this repo's path, and every prognostic field is compared with the reference's
final state at every committed point — no outlier allowance.  Fixtures: tests/golden/make_step_strict.py.

  c12k2n6     c12 L79 layout (1,1), k_split=2, n_split=6, 8 non-zero tracers  (the bench's split: covers the it>0
              substeps, riem_solver3 non-last calls, remap with last_step=False, moist_cv with condensates)
  c24L2k2n3   c24 L79 layout (2,2) = 24 subdomains, k_split=2, n_split=3, 8 non-zero tracers (the bench's decomposition)
  c24L2       c24 L79 layout (2,2), k_split=n_split=1 (the c12 dycore_config)
  c12         c12 L79 layout (1,1), k_split=n_split=1

Tolerance per point (reference metric, util/pace/util/testing/comparison.py:6-68): relative <= 1e-10, OR below the
round-off-sensitivity floor the reference calibrated for one c12 step (tests/savepoint/thresholds/fv_dynamics.yaml:
u, v 2.1e-9; w 6.5e-6 with absolute floor 1.5e-12) — see TOL.  `run_case` also returns the ACHIEVED worst relative
error per field (over the points above the absolute floor), which bench.py carries as "parity".
"""
import json
import os
from datetime import timedelta

import numpy as np

from tests import helpers as H

TRACERS = ["qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke"]
TOL = {"default": (1e-10, 1e-13), "u": (2.1e-9, 1e-11), "v": (2.1e-9, 1e-11), "w": (6.5e-6, 1.5e-12), "ua": (2.1e-9, 1e-11),
       "va": (2.1e-9, 1e-11), "omga": (1e-9, 1e-11), "q_con": (1e-9, 1e-15)}
# multi-substep runs compound the one-step round-off sensitivity: k_split * n_split substeps instead of one
CASES = {
    "c12": dict(dir="c12_step", substeps=1),
    "c12k2n6": dict(dir="c12k2n6_step", substeps=12),
    "c24L2": dict(dir="c24L2_step", inputs="c24L2_inputs", substeps=1),
    "c24L2k2n3": dict(dir="c24L2k2n3_step", substeps=6),
    # do_sat_adj = True: the stock baroclinic_c12.yaml, and k_split = 2 with 8 non-zero tracers
    "c12sat": dict(dir="c12sat_step", substeps=1),
    "c12satk2": dict(dir="c12satk2_step", substeps=2),
    # hord_dp = hord_tm = hord_vt = hord_mt = 5 (the other monotonicity switch of xppm.py:47-61 / xtp_u.py), n_split = 2
    "c12hord5": dict(dir="c12hord5_step", substeps=2),
}


def available(case):
    d = os.path.join(H.GOLDEN, CASES[case]["dir"])
    if not os.path.exists(os.path.join(d, "meta.json")):
        return False
    inp = CASES[case].get("inputs") or json.load(open(os.path.join(d, "meta.json"))).get("inputs")
    return inp is None or os.path.exists(os.path.join(H.GOLDEN, inp, "grid_rank0.npz"))


def load(case):
    d = os.path.join(H.GOLDEN, CASES[case]["dir"])
    meta = json.load(open(os.path.join(d, "meta.json")))
    inp = os.path.join(H.GOLDEN, CASES[case].get("inputs") or meta.get("inputs") or CASES[case]["dir"])
    nranks = 6 * meta["layout"] ** 2
    grids = [dict(np.load(os.path.join(inp, f"grid_rank{r}.npz"))) for r in range(nranks)]
    s0 = [dict(np.load(os.path.join(inp, f"state0_rank{r}.npz"))) for r in range(nranks)]
    from pace_b200.fv3core.dycore_state import FIELDS

    for z in s0:  # fields absent from a reduced input dump are identically zero
        shape3 = z["delp"].shape
        for n, (dims, _) in FIELDS.items():
            if n not in z:
                z[n] = np.zeros(shape3 if len(dims) == 3 else shape3[:2])
    if meta.get("fill_tracers"):
        for z in s0:
            for m, n in enumerate(TRACERS[1:], start=1):
                z[n] = z["qvapor"] * (m + 1) * 0.1
    ranks = meta.get("ranks") or [r for r in range(nranks) if os.path.exists(os.path.join(d, f"state1_rank{r}.npz"))]
    s1 = {r: dict(np.load(os.path.join(d, f"state1_rank{r}.npz"))) for r in ranks}
    return meta, grids, s0, s1


def build(meta, grids, s0, dev=None, process_comm=None, checkpointer=None):
    """`process_comm`: several processes (GPUs) share the case; grids / s0 are then sliced to the local subdomains."""
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.dycore_state import DycoreState
    from pace_b200.fv3core.runtime import Runtime
    from pace_b200.fv3core.stencil_factory import GridIndexing, StencilFactory
    from pace_b200.fv3core.stencils.fv_dynamics import DynamicalCore
    from pace_b200.util.grid.helper import DampingCoefficients, GridData

    nx, layout = meta["nx"], meta["layout"]
    comm, qf = H.make_comm(nx, layout, 79, dev, process_comm=process_comm)
    if process_comm is not None:
        grids = [grids[r] for r in comm.local_ranks]
        s0 = [s0[r] for r in comm.local_ranks]
    gd = GridData.from_arrays(qf, grids)
    damp = DampingCoefficients.from_arrays(qf, grids)
    extra = {}
    if meta.get("hord") is not None:
        extra = dict(hord_dp=meta["hord"], hord_tm=meta["hord"], hord_vt=meta["hord"], hord_mt=meta["hord"])
    cfg = baroclinic_config(nx, (layout, layout), n_split=meta.get("n_split", 1), k_split=meta.get("k_split", 1),
                            do_sat_adj=bool(meta.get("do_sat_adj", False)), **extra)
    rt = Runtime(comm, qf, gd, damp, cfg)
    sf = StencilFactory(None, GridIndexing.from_sizer_and_communicator(qf.sizer, comm), rt)
    state = DycoreState.init_from_numpy_arrays(s0, qf)
    dycore = DynamicalCore(comm, gd, sf, qf, damp, cfg, state.phis, state, timedelta(seconds=cfg.dt_atmos),
                           checkpointer=checkpointer)
    return dycore, state


def compare(out, s1, meta, fields=None, first_rank=0):
    """(failures, achieved): achieved[field] = (worst relative error over points above the floor, worst |diff|).
    `out[field][r - first_rank]` is compared with the reference rank r."""
    n = meta["nx"] // meta["layout"]
    levels = meta.get("levels")
    fields = fields or meta.get("fields") or [f for f in next(iter(s1.values())) if f in out]
    failures, achieved = [], {}
    for name in fields:
        rel, floor = TOL.get(name, TOL["default"])
        worst_rel = worst_abs = 0.0
        for r, z in s1.items():
            a, b = out[name][r - first_rank], z[name]
            if a.ndim == 3:
                nk = 80 if name in ("pe", "peln", "pk") else 79
                ii = slice(3, 3 + n + (1 if name == "v" else 0))
                jj = slice(3, 3 + n + (1 if name == "u" else 0))
                if levels is not None:
                    lv = [k for k in levels if k < nk]
                    a, b = a[ii, jj][:, :, lv], b[ii, jj][:, :, :len(lv)]
                else:
                    a, b = a[ii, jj, :nk], b[ii, jj, :nk]
            else:
                a, b = a[3:3 + n, 3:3 + n], b[3:3 + n, 3:3 + n]
            m = H.ref_metric(a, b)
            d = np.abs(a - b)
            above = d > floor
            if above.any():
                worst_rel = max(worst_rel, float(m[above].max()))
            worst_abs = max(worst_abs, float(d.max()))
            bad = (m > rel) & above
            if bad.any():
                failures.append(f"{name} rank {r}: {int(bad.sum())} pts, worst rel {m[bad].max():.2e}, worst abs {d[bad].max():.2e}")
        achieved[name] = (worst_rel, worst_abs)
    return failures, achieved


def run_case(case, dev=None):
    meta, grids, s0, s1 = load(case)
    dycore, state = build(meta, grids, s0, dev)
    dycore.step_dynamics(state)
    H.sync()
    out = state.as_numpy()
    failures, achieved = compare(out, s1, meta)
    # total dry mass is conserved to round-off (north_star)
    n = meta["nx"] // meta["layout"]
    c = slice(3, 3 + n)
    m0 = sum((s0[r]["delp"][c, c, :79] * grids[r]["area"][c, c, None]).sum() for r in range(len(grids)))
    m1 = sum((out["delp"][r][c, c, :79] * grids[r]["area"][c, c, None]).sum() for r in range(len(grids)))
    return failures, achieved, abs(m1 - m0) / m0
