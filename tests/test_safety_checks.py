"""SafetyChecker (pace_b200/util/safety_checks.py over fv3_field_check) against numpy and against the behaviour of the
reference's driver/pace/driver/safety_checks.py:24-110: min / max over the view or the whole storage, NaN detection,
the exceptions it raises, and the reference's own test cases (driver/tests/test_safety_checks.py: a variable inside its
bounds passes, one outside raises RuntimeError, a doubly registered one raises NotImplementedError)."""
import numpy as np
import pytest

from tests import helpers as H


def _case(dev):
    got = H.load_case("c12", (0, 0), dev)
    if got is None:
        pytest.skip("c12 golden case not available")
    return got


def _check(dev):
    from pace_b200.util import safety_checks as S

    comm, qf, rt, sf = _case(dev)
    rng = np.random.default_rng(7)
    q3, qi, q2 = qf.zeros(H.D3, "unknown"), qf.zeros(H.D3I, "unknown"), qf.zeros(H.D2, "unknown")
    for q in (q3, qi, q2):
        q.data.copy_(H.torch.as_tensor(rng.normal(size=tuple(q.data.shape)) * 10.0).to(q.data.device))
    for q in (q3, qi, q2):
        a = q.data.cpu().numpy()
        sl = (slice(None),) + tuple(slice(o, o + e) for o, e in zip(q.origin, q.extent))
        assert S.field_min_max_nan(rt, q, False) == (a.min(), a.max(), 0)
        assert S.field_min_max_nan(rt, q, True) == (a[sl].min(), a[sl].max(), 0)
    # NaNs are counted and ignored by min / max; signed zeros and infinities keep their order
    o = q3.origin
    q3.data[0, o[0] + 1, o[1] + 2, 5] = float("nan")
    q3.data[1, 0, 0, 0] = float("nan")           # in the halo of the second subdomain
    q3.data[1, o[0], o[1], 0] = float("-inf")
    a = q3.data.cpu().numpy()
    sl = (slice(None),) + tuple(slice(p, p + e) for p, e in zip(q3.origin, q3.extent))
    assert S.field_min_max_nan(rt, q3, False) == (float("-inf"), np.nanmax(a), 2)
    assert S.field_min_max_nan(rt, q3, True) == (float("-inf"), np.nanmax(a[sl]), 1)

    class State:
        pass

    st = State()
    st.delp, st.pt = qf.zeros(H.D3, "Pa"), qf.zeros(H.D3, "K")
    st.delp.data.fill_(500.0)
    st.pt.data.fill_(280.0)
    S.SafetyChecker.clear_all_checks()
    S.SafetyChecker.register_variable("delp", minimum_value=1e-3, maximum_value=1e5, compute_domain_only=True)
    S.SafetyChecker.register_variable("pt", minimum_value=150.0, maximum_value=400.0)
    with pytest.raises(NotImplementedError):
        S.SafetyChecker.register_variable("pt", 0.0, 1.0)
    chk = S.SafetyChecker(rt)
    chk.check_state(st)                           # inside the bounds: passes
    st.delp.data[0, 0, 0, 0] = -1.0               # negative delp in the halo: the compute-domain check does not see it
    chk.check_state(st)
    st.delp.data[1, o[0] + 3, o[1] + 3, 7] = -1.0
    with pytest.raises(RuntimeError, match="delp is outside of its specified bounds"):
        chk.check_state(st)
    st.delp.data[1, o[0] + 3, o[1] + 3, 7] = 500.0
    st.pt.data[0, 1, 1, 1] = 500.0                # whole-storage check: a halo value counts
    with pytest.raises(RuntimeError, match="pt is outside of its specified bounds"):
        chk.check_state(st)
    st.pt.data[0, 1, 1, 1] = 280.0
    st.pt.data[0, o[0], o[1], 3] = float("nan")
    with pytest.raises(RuntimeError, match="pt contains a NaN value"):
        chk.check_state(st)
    S.SafetyChecker.clear_all_checks()
    S.SafetyChecker.register_variable("not_there", 0.0, 1.0)
    with pytest.raises(NotImplementedError, match="not in the state"):
        chk.check_state(st)
    S.SafetyChecker.clear_all_checks()


def test_safety_checks_hostsim(device):
    if device != "cpu":
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _check(device)


@pytest.mark.gpu
def test_safety_checks_gpu(device):
    _check(device)
