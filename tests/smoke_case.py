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
