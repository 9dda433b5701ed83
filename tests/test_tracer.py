"""TracerAdvection: the fused per-sub-cycle kernel (fv3_tracer_subcycle) must give bit-identical results to the
reference-shaped sequence apply_mass_flux -> fvtp2d(hord_tr) -> apply_tracer_flux -> swap_dp per tracer
(tracer_2d_1l.py:341-392), on a full c12 baroclinic step with all 8 tracers non-zero."""
from datetime import timedelta

import numpy as np
import pytest

from tests import helpers as H


def _step(fused, dev):
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.initialization import baroclinic
    from pace_b200.fv3core.runtime import Runtime
    from pace_b200.fv3core.stencil_factory import GridIndexing, StencilFactory
    from pace_b200.fv3core.stencils.fv_dynamics import DynamicalCore
    from pace_b200.util.grid.helper import DampingCoefficients, GridData

    comm, qf = H.make_comm(12, 1, 79, dev)
    gd = GridData.new_from_generation(qf, comm)
    damp = DampingCoefficients.new_from_generation(qf, gd)
    cfg = baroclinic_config(12, (1, 1), n_split=1, k_split=1)
    rt = Runtime(comm, qf, gd, damp, cfg)
    sf = StencilFactory(None, GridIndexing.from_sizer_and_communicator(qf.sizer, comm), rt)
    state = baroclinic.init_baroclinic_state(gd, qf, False, False, True, comm, fill_all_tracers=True)
    dycore = DynamicalCore(comm, gd, sf, qf, damp, cfg, state.phis, state, timedelta(seconds=cfg.dt_atmos))
    dycore.tracer_advection._fused = fused
    dycore.step_dynamics(state)
    H.sync()
    return state.as_numpy()


def _check(dev):
    a, b = _step(True, dev), _step(False, dev)
    for name in ("qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke", "delp", "pt", "u", "w"):
        np.testing.assert_array_equal(a[name], b[name], err_msg=name)
    assert np.abs(a["qsgs_tke"][:, 3:15, 3:15, :79]).max() > 0


def test_fused_subcycle_bit_identical_hostsim(device):
    if device != "cpu":
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _check(device)


@pytest.mark.gpu
def test_fused_subcycle_bit_identical_gpu(device):
    _check(device)
