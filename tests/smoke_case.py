"""The smoke invocation: stage kernels on the committed c12 golden inputs, checked against the numpy oracle."""
import numpy as np

from oracle import riem_solver as O
from tests import helpers as H


def run(verbose=False):
    from pace_b200.fv3core.stencils.riem_solver_c import NonhydrostaticVerticalSolverCGrid

    d = H.load_stage("c12", 0, "Riem_Solver_C#0")
    comm, qf, rt, sf = H.load_case("c12", (0,))
    q = {k[3:]: H.to_q(qf, [v]) for k, v in d.items() if k.startswith("in.") and v.ndim >= 2}
    NonhydrostaticVerticalSolverCGrid(sf, qf, 0.05)(
        float(d["in.dt2"]), q["cappa"], float(d["in.ptop"]), q["hs"], q["ws"], q["ptc"], q["q_con"], q["delpc"],
        q["gz"], q["pef"], q["w3"])
    H.sync()
    gz, pef = d["in.gz"].copy(), d["in.pef"].copy()
    O.riem_solver_c(float(d["in.dt2"]), d["in.cappa"], float(d["in.ptop"]), d["in.hs"], d["in.ws"], d["in.ptc"],
                    d["in.q_con"], d["in.delpc"], gz, pef, d["in.w3"], 0.05, 12, 12, 79)
    H.assert_close(q["gz"].numpy()[0], gz, 1e-12, name="gz")
    H.assert_close(q["pef"].numpy()[0], pef, 1e-12, name="pef")
    if verbose:
        print("riem_solver_c: CUDA == oracle within 1e-12")
    # the hottest stage and the vertical remap, same way: native call vs the numpy restatement on the committed inputs
    from oracle.indexing import Idx
    from tests.stage_specs import NX, NZ, SPECS
    from tests.test_stages import _grid, run_native

    for name in ("d_sw@s2", "remapping"):
        spec = SPECS[name]
        d = H.load_stage(spec.case, 0, spec.golden)
        q = run_native(spec, d)
        a = {k[3:]: v.copy() for k, v in d.items() if k.startswith("in.")}
        spec.oracle(Idx(NX, NX, NZ), _grid(spec.case), a)
        for n in spec.outputs:
            reg = spec.regions.get(n, (slice(None), slice(None)))
            H.assert_close(q[n].numpy()[0][reg], a[n][reg], max(spec.tols.get(n, spec.tol), 1e-11), max(spec.near_zero, 1e-13),
                           name=f"{name}.{n}")
        if verbose:
            print(f"{name}: CUDA == oracle")
