"""In-flight savepoints of a full step against the reference's: the benchmark's split (k_split=2, n_split=6, 8 non-zero
tracers) at c12, driven with a recording checkpointer.  Every time `step_dynamics` reaches one of the reference's
savepoints (fv_dynamics.py:340-422 Remapping-In/-Out, Tracer2D1L-In/-Out; dyn_core.py:608-668 D_SW-Out) the variables it
hands over are compared, on rank 0, with what the unmodified reference held at the same call (stage captures of
oracle/refshim, reduced by tests/golden/make_step_strict.py:c12k2n6_checkpoints).  The end-state tests prove the step;
this one proves that the state BETWEEN the stages is the reference's too — what a savepoint-by-savepoint translate test
of the reference would look at.  Tolerances are those of tests/step_cases.py (the in-flight fields use the default).
"""
import os

import numpy as np
import pytest

from tests import helpers as H
from tests import step_cases as SC

FIXTURE = os.path.join(H.GOLDEN, "c12k2n6_step", "checkpoints_rank0.npz")
# savepoint variable -> the field whose tolerance applies
ALIAS = dict(ud="u", vd="v", wd="w", uad="ua", vad="va", ptd="pt", delpd="delp")
X_FACES = ("v", "vd", "mfxd", "xfxd", "cxd")  # one more point along i
Y_FACES = ("u", "ud", "mfyd", "yfxd", "cyd")  # one more point along j
FLUXES = ("mfxd", "mfyd", "xfxd", "yfxd", "cxd", "cyd")
N = 12


class Recorder:
    """checkpointer(savepoint_name, **quantities): compares on the fly, keeps (key, worst rel, worst abs)."""

    def __init__(self, ref, state):
        self.ref, self.state = ref, state
        self.levels = [int(k) for k in ref["levels"]]
        self.calls, self.seen, self.failures, self.worst = {}, set(), [], {}

    def __call__(self, name, **variables):
        n = self.calls.get(name, 0)
        self.calls[name] = n + 1
        if name == "Tracer2D1L-Out":
            variables = dict(variables, **self.state.tracers)
        for var, q in variables.items():
            key = f"{name}#{n}/{var}"
            if key not in self.ref:
                continue
            ni, nj = N + (var in X_FACES), N + (var in Y_FACES)
            b = self.ref[key][:ni, :nj]
            a = q.numpy()[0][3:3 + ni, 3:3 + nj]
            if a.ndim == 3:
                a = a[:, :, self.levels[:b.shape[2]]]
            self.seen.add(key)
            rel, floor = SC.TOL.get(ALIAS.get(var, var), SC.TOL["default"])
            m, d = H.ref_metric(a, b), np.abs(a - b)
            if var in FLUXES:  # |values| up to 1e12 that pass through zero: the absolute floor scales with the field
                floor = max(floor, 1e-12 * float(np.abs(b).max()))
            above = d > floor
            self.worst[key] = (float(m[above].max()) if above.any() else 0.0, float(d.max()) if d.size else 0.0)
            bad = (m > rel) & above
            if bad.any():
                self.failures.append(f"{key}: {int(bad.sum())} pts, worst rel {m[bad].max():.2e}, worst abs {d[bad].max():.2e}")


def _run(dev):
    if not (SC.available("c12k2n6") and os.path.exists(FIXTURE)):
        pytest.skip("c12k2n6 checkpoint fixture not available")
    ref = dict(np.load(FIXTURE))
    meta, grids, s0, _ = SC.load("c12k2n6")
    holder = {}
    rec = lambda name, **kw: holder["rec"](name, **kw)
    dycore, state = SC.build(meta, grids, s0, dev, checkpointer=rec)
    holder["rec"] = Recorder(ref, state)
    dycore.step_dynamics(state)
    H.sync()
    r = holder["rec"]
    missing = sorted(k for k in ref if k != "levels" and k not in r.seen)
    assert not missing, f"savepoints never reached: {missing[:8]}"
    assert r.calls["D_SW-Out"] == 12 and r.calls["Remapping-In"] == 2 and r.calls["Tracer2D1L-Out"] == 2
    print("\nworst per savepoint:", {sp: max(v[0] for k, v in r.worst.items() if k.startswith(sp)) for sp in
                                      ("D_SW-Out", "Tracer2D1L-In", "Tracer2D1L-Out", "Remapping-In", "Remapping-Out")})
    assert not r.failures, "\n".join(r.failures)


def test_savepoints_match_reference_hostsim(device):
    if device != "cpu":
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _run(device)


@pytest.mark.gpu
def test_savepoints_match_reference_gpu(device):
    _run(device)
