"""Halo exchange: device gather kernel vs the host table application, all field staggerings and layouts."""
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
@pytest.mark.parametrize("layout", [1, 2])
def test_halo_update_gpu(layout):
    _run(layout, C3, None, 3)
    _run(layout, B3, None, 2)
    _run(layout, U3, V3, 3)
    _run(layout, V3, U3, 3)
    _run(layout, U3, V3, 0, mode="interface")


def test_zero_halo_points_is_an_error():
    comm, qf = H.make_comm(8, 1, 3)
    with pytest.raises(ValueError):
        comm.get_scalar_halo_updater([qf.get_quantity_halo_spec(C3, 0)])
