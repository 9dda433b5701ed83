"""AdjustNegativeTracerMixingRatio (fv3_neg_adj3) against the UNMODIFIED reference.

tests/golden/neg_adj3/case0.npz: inputs / outputs of the reference's neg_adj3.py (numpy backend) on seeded synthetic
12 x 12 x 79 fields in which 30 % of the values of every water species are negative, plus whole columns in deficit and
negative top / bottom levels (generator: oracle/refshim/gen_neg_adj.py).  The reference leaves non-negative input
bit-identical (the generator checks it); so must we, including on the analytic baroclinic state of the full-step test.
Tolerance: 1e-13 relative (reference metric), 1e-25 absolute floor for the round-off residues the borrowing leaves.
"""
import os

import numpy as np
import pytest

from tests import helpers as H

NAMES = ["qvapor", "qliquid", "qrain", "qsnow", "qice", "qgraupel", "qcld", "pt", "delp"]


def _run(dev, fields):
    got = H.load_case("c12", (0,), dev)
    if got is None:
        pytest.skip("c12 golden case not available")
    comm, qf, rt, sf = got
    from pace_b200.fv3core.stencils.neg_adj3 import AdjustNegativeTracerMixingRatio

    qs = {}
    for n in NAMES:
        q = qf.zeros(H.D3, "unknown")
        q.data[0, 3:15, 3:15, :79] = H.torch.as_tensor(fields[n]).to(q.data.device)
        qs[n] = q
    AdjustNegativeTracerMixingRatio(sf, qf, check_negative=False, hydrostatic=False)(*[qs[n] for n in NAMES])
    H.sync()
    return {n: qs[n].data[0, 3:15, 3:15, :79].cpu().numpy() for n in NAMES}


def _check(dev):
    p = os.path.join(H.GOLDEN, "neg_adj3", "case0.npz")
    z = np.load(p)
    out = _run(dev, {n: z["in." + n] for n in NAMES})
    for n in NAMES:
        H.assert_close(out[n], z["out." + n], max_error=1e-13, near_zero=1e-25, name=n)
    assert (z["out.qvapor"] != z["in.qvapor"]).sum() > 1000  # the case does exercise the borrowing
    # non-negative input: untouched, bit for bit
    pos = {n: np.abs(z["in." + n]) for n in NAMES}
    out = _run(dev, pos)
    for n in NAMES:
        np.testing.assert_array_equal(out[n], pos[n], err_msg=n)


def test_oracle_neg_adj3_matches_reference():
    """The Python restatement (oracle/negative_water.py) against the reference's own output."""
    from oracle import negative_water as O

    z = np.load(os.path.join(H.GOLDEN, "neg_adj3", "case0.npz"))
    a = {n: z["in." + n].copy() for n in NAMES}
    O.neg_adj3(*[a[n] for n in NAMES])
    for n in NAMES:
        H.assert_close(a[n], z["out." + n], max_error=1e-13, near_zero=1e-25, name=n)


def test_neg_adj3_hostsim(device):
    if device != "cpu":
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _check(device)


@pytest.mark.gpu
def test_neg_adj3_gpu(device):
    _check(device)
