"""Strip decomposition of the plane-resident kernels (pace_b200/csrc/plane.h).

At c12 a plane fits in shared memory as ONE strip, so the golden-vector tests would never run the multi-strip logic
(row clipping of every phase, halo rows recomputed by two strips, parked boundary rows of the in-place tracer update).
`FV3_FORCE_STRIPS=n` makes the launcher cut every plane into n strips; it is read once per process, hence the
subprocesses.  The same golden comparisons (reference outputs, tests/golden/) must hold for 2 strips (6 rows each) and
3 strips (4 rows each: thinner than the 2 x 3 halo rows a strip reads from its neighbours)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANE_TESTS = ["tests/test_stages.py", "tests/test_tracer.py", "tests/test_dycore_step.py"]


def _run(n_strips, marker):
    env = dict(os.environ, FV3_FORCE_STRIPS=str(n_strips))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", marker, "-p", "no:cacheprovider"] + PLANE_TESTS,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, f"{n_strips} strips:\n{r.stdout[-3000:]}\n{r.stderr[-2000:]}"
    assert " passed" in r.stdout


@pytest.mark.parametrize("n_strips", [2, 3])
def test_forced_strips_hostsim(device, n_strips):
    if device != "cpu":
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _run(n_strips, "not gpu")


@pytest.mark.gpu
def test_forced_strips_gpu():
    _run(2, "gpu")


_C48_SCRIPT = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from oracle import hostsim
hostsim.install(openmp=True)
from tests.test_c48_step import _build
dycore, state = _build("cpu")
dycore.step_dynamics(state)
out = state.as_numpy()
h = hashlib.sha1()
for n in ("u", "v", "w", "delp", "pt", "delz", "qvapor"):
    h.update(np.ascontiguousarray(out[n]).tobytes())
print("DIGEST", h.hexdigest())
"""


def _c48_digest(n_strips):
    env = dict(os.environ)
    env.pop("FV3_FORCE_STRIPS", None)
    if n_strips:
        env["FV3_FORCE_STRIPS"] = str(n_strips)
    r = subprocess.run([sys.executable, "-c", _C48_SCRIPT % ROOT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-3000:]
    return [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0]


def test_c48_strip_count_does_not_change_a_bit(device):
    """48 x 48 subdomains: the launcher's own choice (2 strips of 24 rows for the 5-plane kernels), one strip and 5
    strips of 10 rows give the same timestep bit for bit (host simulation of the kernel sources)."""
    if device != "cpu":
        pytest.skip("host simulation is exercised on CPU-only boxes")
    auto, one, five = _c48_digest(0), _c48_digest(1), _c48_digest(5)
    assert auto == one == five


_SMALL_SCRIPT = r"""
import hashlib, sys
from datetime import timedelta
import numpy as np
sys.path.insert(0, %r)
from oracle import hostsim
hostsim.install(openmp=True)
import bench
dycore, state, comm, rt, gd = bench.build_dycore(%d, 1, 79, 1, 2, "cpu", all_tracers=True)
dycore.step_dynamics(state)
out = state.as_numpy()
h = hashlib.sha1()
for n in ("u", "v", "w", "delp", "pt", "delz", "qvapor", "qliquid", "qsgs_tke"):
    h.update(np.ascontiguousarray(out[n]).tobytes())
print("DIGEST", h.hexdigest())
"""


def _small_digest(nx, n_strips):
    env = dict(os.environ)
    env.pop("FV3_FORCE_STRIPS", None)
    if n_strips:
        env["FV3_FORCE_STRIPS"] = str(n_strips)
    r = subprocess.run([sys.executable, "-c", _SMALL_SCRIPT % (ROOT, nx)], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-3000:]
    return [l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0]


@pytest.mark.parametrize("nx,n_strips", [(13, 4), (17, 4)])
def test_last_strip_shorter_than_the_halo(device, nx, n_strips):
    """ny % rows_per_strip in {1, 2}: c13 in 4 strips = 4 + 4 + 4 + 1 rows, c17 in 4 strips = 5 + 5 + 5 + 2 rows: the in-place
    tracer update parks the rows a neighbouring strip reads as halo and copies them back afterwards; rows beyond the
    compute domain were never parked and must not be copied (8 non-zero tracers, two substeps, one timestep)."""
    if device != "cpu":
        pytest.skip("host simulation is exercised on CPU-only boxes")
    assert _small_digest(nx, 1) == _small_digest(nx, n_strips)
