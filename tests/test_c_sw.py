"""c_sw (+ d2a2c_vect): oracle vs reference golden; native kernels vs golden."""
import numpy as np
import pytest

from oracle import c_sw as O
from oracle.indexing import Idx
from tests import helpers as H

CASE = "c12"
OUT = ("delp", "pt", "w", "uc", "vc", "ua", "va", "ut", "vt", "divgd", "omga", "delpc", "ptc")


def _golden():
    d = H.load_stage(CASE, 0, "C_SW#0")
    if d is None:
        pytest.skip("golden vectors not available")
    return d


def test_oracle_c_sw_matches_reference():
    d = _golden()
    g = dict(np.load(H.golden_path(CASE, "grid_rank0.npz")))
    a = {k[3:]: v.copy() for k, v in d.items() if k.startswith("in.")}
    delpc, ptc = O.c_sw(Idx(12, 12, 79), g, a["delp"], a["pt"], a["u"], a["v"], a["w"], a["uc"], a["vc"], a["ua"],
                        a["va"], a["ut"], a["vt"], a["divgd"], a["omga"], float(a["dt2"]))
    a["delpc"], a["ptc"] = delpc, ptc
    for n in OUT:
        H.assert_close(a[n], d["out." + n], 1e-14, name=n)


def _run_native(d):
    from pace_b200.fv3core.stencils.c_sw import CGridShallowWaterDynamics

    comm, qf, rt, sf = H.load_case(CASE, (0,))
    q = {k[3:]: H.to_q(qf, [v]) for k, v in d.items() if k.startswith("in.") and v.ndim >= 2}
    csw = CGridShallowWaterDynamics(sf, qf, rt.grid_data, False, 0, 3)
    delpc, ptc = csw(q["delp"], q["pt"], q["u"], q["v"], q["w"], q["uc"], q["vc"], q["ua"], q["va"], q["ut"], q["vt"],
                     q["divgd"], q["omga"], float(d["in.dt2"]))
    H.sync()
    q["delpc"], q["ptc"] = delpc, ptc
    return q


def _check(d, q):
    for n in OUT:
        H.assert_close(q[n].numpy()[0], d["out." + n], 1e-14, name=n)


def test_native_c_sw_hostsim():
    d = _golden()
    _check(d, _run_native(d))


@pytest.mark.gpu
def test_native_c_sw_gpu():
    d = _golden()
    _check(d, _run_native(d))
