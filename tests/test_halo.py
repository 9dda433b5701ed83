"""Halo exchange through the communicator API (device gather / pack / unpack kernels): (1) against the arrays the
REFERENCE's own halo updater produces from index-encoded fields (tests/golden/topology/halo_known_answers.npz, made by
tests/golden/make_halo_known_answers.py; layouts 1-3, every staggering of the hot path), (2) random fields against
the host application of the same table (kernel vs table consistency for more sizes)."""
import numpy as np
import pytest
import torch

from pace_b200 import constants as c
from pace_b200.util import topology as T
from tests import helpers as H

C3 = (c.X_DIM, c.Y_DIM, c.Z_DIM)
B3 = (c.X_INTERFACE_DIM, c.Y_INTERFACE_DIM, c.Z_DIM)
U3 = (c.X_DIM, c.Y_INTERFACE_DIM, c.Z_DIM)
V3 = (c.X_INTERFACE_DIM, c.Y_DIM, c.Z_DIM)
ZI = (c.X_DIM, c.Y_DIM, c.Z_INTERFACE_DIM)


def _stag(d):
    return (0 if d[0] == c.X_INTERFACE_DIM else 1, 0 if d[1] == c.Y_INTERFACE_DIM else 1)


def _run(layout, dims_x, dims_y, n_halo, mode="halo"):
    nz = 5
    comm, qf = H.make_comm(8 * layout, layout, nz)
    rng = np.random.default_rng(1)
    qx = qf.zeros(dims_x, "m")
    qx.set_from_numpy(rng.standard_normal(qx.shape))
    arrs_x = [a.copy() for a in qx.numpy()]
    qy = arrs_y = None
    if dims_y is not None:
        qy = qf.zeros(dims_y, "m")
        qy.set_from_numpy(rng.standard_normal(qy.shape))
        arrs_y = [a.copy() for a in qy.numpy()]
    table = T.build_halo_table(comm.decomposition, n_halo, _stag(dims_x), _stag(dims_y) if dims_y else None, mode=mode)
    nlev = nz + 1 if dims_x[2] == c.Z_INTERFACE_DIM else nz
    ex = [a[:, :, :nlev] for a in arrs_x]
    ey = [a[:, :, :nlev] for a in arrs_y] if arrs_y else None
    T.apply_table_numpy(table, ex, ey)
    if mode == "interface":
        comm.synchronize_vector_interfaces(qx, qy)
    elif dims_y is None:
        up = comm.get_scalar_halo_updater([qf.get_quantity_halo_spec(dims_x, n_halo)])
        up.start([qx])
        with pytest.raises(RuntimeError):
            up.start([qx])
        up.wait()
    else:
        comm.vector_halo_update(qx, qy, n_halo)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    np.testing.assert_array_equal(qx.numpy(), np.stack(arrs_x))
    if qy is not None:
        np.testing.assert_array_equal(qy.numpy(), np.stack(arrs_y))


@pytest.mark.parametrize("layout", [1, 2])
@pytest.mark.parametrize("dims,n_halo", [(C3, 3), (C3, 2), (B3, 3), (ZI, 3)])
def test_scalar_halo_update_hostsim(layout, dims, n_halo):
    _run(layout, dims, None, n_halo)


@pytest.mark.parametrize("layout", [1, 2])
@pytest.mark.parametrize("dx,dy", [(U3, V3), (V3, U3)])
def test_vector_halo_update_hostsim(layout, dx, dy):
    _run(layout, dx, dy, 3)


@pytest.mark.parametrize("layout", [1, 2])
def test_synchronize_vector_interfaces_hostsim(layout):
    _run(layout, U3, V3, 0, mode="interface")


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [1, 2, 3])
def test_halo_update_gpu(layout):
    _run(layout, C3, None, 3)
    _run(layout, B3, None, 2)
    _run(layout, U3, V3, 3)
    _run(layout, V3, U3, 3)
    _run(layout, U3, V3, 0, mode="interface")


_DIMS = {"x": c.X_DIM, "y": c.Y_DIM, "x_interface": c.X_INTERFACE_DIM, "y_interface": c.Y_INTERFACE_DIM}


def _known_answer(layout, case):
    """Index-encoded fields through comm.halo_update / vector_halo_update / synchronize_vector_interfaces on the
    device, compared bit for bit with what the reference's halo updater returned for the same fields."""
    import os

    from tests.test_topology import CASE_DIMS, TOPO, _encode

    z = np.load(os.path.join(TOPO, "halo_known_answers.npz"))
    dx, dy, nh, mode = CASE_DIMS[case]
    zdim = c.Z_INTERFACE_DIM if case == "scalar_zi_h3" else c.Z_DIM
    ref_x = z[f"L{layout}.{case}.x"]
    ref_y = z[f"L{layout}.{case}.y"] if dy is not None else None
    nz = 2
    comm, qf = H.make_comm(4 * layout, layout, nz)
    qx = qf.zeros((_DIMS[dx[0]], _DIMS[dx[1]], zdim), "m")
    n = ref_x.shape[0]
    qx.set_from_numpy(np.stack([np.repeat(_encode(ref_x.shape[1:], r, 1.0)[:, :, None], qx.shape[3], 2) for r in range(n)]))
    qy = None
    if dy is not None:
        qy = qf.zeros((_DIMS[dy[0]], _DIMS[dy[1]], c.Z_DIM), "m")
        qy.set_from_numpy(np.stack([np.repeat(_encode(ref_y.shape[1:], r, 0.5)[:, :, None], qy.shape[3], 2) for r in range(n)]))
    if mode == "interface":
        comm.synchronize_vector_interfaces(qx, qy)
    elif qy is None:
        comm.halo_update(qx, nh)
    else:
        comm.vector_halo_update(qx, qy, nh)
    H.sync()
    nlev = nz + 1 if zdim == c.Z_INTERFACE_DIM else nz
    for k in range(nlev):
        np.testing.assert_array_equal(qx.numpy()[:, :, :, k], ref_x)
        if qy is not None:
            np.testing.assert_array_equal(qy.numpy()[:, :, :, k], ref_y)


_KA_CASES = ["scalar_cell_h3", "scalar_cell_h2", "scalar_corner_h3", "scalar_zi_h3", "vector_dgrid_h3", "vector_cgrid_h3",
             "vector_agrid_h3", "sync_interfaces_dgrid"]


@pytest.mark.parametrize("layout", [1, 2, 3])
@pytest.mark.parametrize("case", _KA_CASES)
def test_halo_update_matches_reference_hostsim(layout, case):
    if torch.cuda.is_available():
        pytest.skip("host simulation is exercised on CPU-only boxes")
    _known_answer(layout, case)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [1, 2, 3])
@pytest.mark.parametrize("case", _KA_CASES)
def test_halo_update_matches_reference_gpu(layout, case):
    _known_answer(layout, case)


def test_zero_halo_points_is_an_error():
    comm, qf = H.make_comm(8, 1, 3)
    with pytest.raises(ValueError):
        comm.get_scalar_halo_updater([qf.get_quantity_halo_spec(C3, 0)])
