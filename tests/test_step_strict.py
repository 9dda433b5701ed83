"""STRICT one-timestep parity with the reference, including the benchmarked configuration (see tests/step_cases.py):
reference grid + initial state in, every prognostic field and all 8 tracers compared at every committed point, NO
outlier allowance; the achieved worst error per field is printed (pytest -s) and asserted."""
import pytest
import torch

from tests import step_cases as S


def _run(case):
    if not S.available(case):
        pytest.skip(f"{case}: fixtures not present")
    failures, achieved, dmass = S.run_case(case)
    print(f"\n[{case}] achieved worst error per field (relative above the floor, absolute):")
    for name, (rel, ab) in achieved.items():
        print(f"   {name:10s} rel {rel:9.2e}   abs {ab:9.2e}")
    print(f"   total dry mass relative change {dmass:.2e}")
    assert not failures, "\n".join(failures)
    assert dmass < 1e-13


@pytest.mark.parametrize("case", sorted(S.CASES))
def test_step_strict_hostsim(case):
    if torch.cuda.is_available():
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _run(case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(S.CASES))
def test_step_strict_gpu(case):
    _run(case)
