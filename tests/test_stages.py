"""Per-stage parity: oracle vs reference golden vectors, native kernels vs golden, for every registered stage.

Golden vectors were dumped from the UNMODIFIED reference (numpy backend) at its own call sites on the c12
baroclinic case (oracle/refshim/gen_golden.py, tests/golden/make_committed.py).  Tolerances use the reference's
own metric (util/pace/util/testing/comparison.py:6-68).
"""
import numpy as np
import pytest

from oracle.indexing import Idx
from tests import helpers as H
from tests.stage_specs import NX, NZ, SPECS

def _golden(spec):
    d = H.load_stage(spec.case, 0, spec.golden)
    if d is None:
        pytest.skip(f"golden vectors for {spec.golden} not available")
    return d


def _grid(case="c12"):
    return dict(np.load(H.golden_path(case, "grid_rank0.npz")))


@pytest.mark.parametrize("name", sorted(SPECS))
def test_oracle_matches_reference(name):
    spec = SPECS[name]
    d = _golden(spec)
    a = {k[3:]: v.copy() for k, v in d.items() if k.startswith("in.")}
    spec.oracle(Idx(NX, NX, NZ), _grid(spec.case), a)
    for n in spec.outputs:
        reg = spec.regions.get(n, (slice(None), slice(None)))
        H.assert_close(a[n][reg], d["out." + n][reg], spec.oracle_tol, spec.near_zero, name=f"{name}.{n}")


def run_native(spec, d):
    comm, qf, rt, sf = H.load_case(spec.case, (0,))
    q = {k[3:]: H.to_q(qf, [v]) for k, v in d.items() if k.startswith("in.") and v.ndim >= 2}
    spec.native(sf, qf, rt, q, d)
    H.sync()
    return q


def _check_native(spec):
    d = _golden(spec)
    q = run_native(spec, d)
    if spec.expected is not None:
        d = spec.expected(d)
    for n in spec.outputs:
        reg = spec.regions.get(n, (slice(None), slice(None)))
        H.assert_close(q[n].numpy()[0][reg], d["out." + n][reg], spec.tols.get(n, spec.tol), spec.near_zero,
                       name=f"{spec.name}.{n}")
    # inputs that the reference leaves untouched must be untouched here as well
    for k, v in (d.items() if spec.check_untouched else ()):
        n = k[3:]
        if k.startswith("in.") and v.ndim >= 2 and n not in spec.outputs and np.array_equal(v, d["out." + n], equal_nan=True):
            np.testing.assert_array_equal(q[n].numpy()[0], v, err_msg=f"{spec.name}: input {n} was modified")


@pytest.mark.parametrize("name", sorted(SPECS))
def test_native_hostsim_matches_reference(name):
    import torch

    if torch.cuda.is_available():
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _check_native(SPECS[name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SPECS))
def test_native_gpu_matches_reference(name):
    _check_native(SPECS[name])
