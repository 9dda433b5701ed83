"""SatAdjust3d (fv3_sat_adjust) against the UNMODIFIED reference: the two SatAdjust3d calls of a
k_split = 2 c12 step with 8 non-zero tracers (tests/golden/c12satk2_step/stage_rank0: call #0 has last_step = False,
call #1 last_step = True and the cloud-fraction branch; generator oracle/refshim/gen_golden.py --do-sat-adj
--fill-tracers, reduced by tests/golden/make_step_strict.py).  Both the numpy restatement (oracle/) and the CUDA kernel
are held to 1e-11 relative (the kernel differs from numpy by the last ulps of exp / log only); the full-step cases with
do_sat_adj are in tests/test_step_strict.py (c12sat = the stock baroclinic_c12.yaml)."""
import os

import numpy as np
import pytest
import torch

from tests import helpers as H

BASE = os.path.join(H.GOLDEN, "c12satk2_step", "stage_rank0")
FIELDS = ["qvapor", "qliquid", "qice", "qrain", "qsnow", "qgraupel", "qcld", "delp", "delz", "q_con", "pt", "pkz", "cappa", "te"]
OUT = ["qvapor", "qliquid", "qice", "qrain", "qsnow", "qgraupel", "qcld", "q_con", "pt", "pkz", "cappa"]
C = slice(3, 15)


def _load(n):
    p = os.path.join(BASE, f"SatAdjust3d#{n}.npz")
    if not os.path.exists(p):
        pytest.skip("SatAdjust3d golden not available")
    z = np.load(p)
    d = {k: z[k] for k in z.files}
    for k in list(d):
        if k.startswith("in.") and "out." + k[3:] not in d:
            d["out." + k[3:]] = d[k]
    return d


def _compare(got, d, what):
    for n in OUT:
        H.assert_close(got[n][C, C, :79], d["out." + n][C, C, :79], max_error=1e-11, near_zero=1e-18, name=f"{what} {n}")


@pytest.mark.parametrize("call", [0, 1])
def test_oracle_sat_adjust_matches_reference(call):
    from oracle import saturation_adjustment as O

    d = _load(call)
    grid = np.load(os.path.join(H.GOLDEN, "c12", "grid_rank0.npz"))
    a = {n: d["in." + n][C, C, :79].copy() for n in FIELDS}
    O.sat_adjust(a, grid["area_64"][C, C], d["in.hs"][C, C], float(d["in.r_vir"]), float(d["in.mdt"]), bool(d["in.fast_mp_consv"]),
                 bool(d["in.last_step"]), int(d["in.kmp"]), 79)
    for n in OUT:
        H.assert_close(a[n], d["out." + n][C, C, :79], max_error=1e-11, near_zero=1e-18, name=f"oracle {n}")
    assert bool(d["in.last_step"]) == (call == 1)
    assert (d["out.qliquid"] != d["in.qliquid"]).sum() > 1000   # the case does exercise the phase changes


def _native(call, dev):
    got = H.load_case("c12", (0,), dev, do_sat_adj=True)
    if got is None:
        pytest.skip("c12 golden case not available")
    comm, qf, rt, sf = got
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.stencils.saturation_adjustment import SatAdjust3d

    d = _load(call)
    q = {n: H.to_q(qf, [d["in." + n]]) for n in FIELDS + ["peln"]}
    hs = H.to_q(qf, [d["in.hs"]])
    sa = SatAdjust3d(sf, baroclinic_config(12, do_sat_adj=True).sat_adjust, rt.grid_data.area_64, int(d["in.kmp"]))
    sa(q["te"], q["qvapor"], q["qliquid"], q["qice"], q["qrain"], q["qsnow"], q["qgraupel"], q["qcld"], hs, q["peln"], q["delp"],
       q["delz"], q["q_con"], q["pt"], q["pkz"], q["cappa"], float(d["in.r_vir"]), float(d["in.mdt"]), bool(d["in.fast_mp_consv"]),
       bool(d["in.last_step"]), float(d["in.akap"]), int(d["in.kmp"]))
    H.sync()
    _compare({n: q[n].numpy()[0] for n in OUT}, d, "kernel")


@pytest.mark.parametrize("call", [0, 1])
def test_sat_adjust_hostsim(call):
    if torch.cuda.is_available():
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _native(call, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("call", [0, 1])
def test_sat_adjust_gpu(call):
    _native(call, "cuda")
