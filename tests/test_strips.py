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
