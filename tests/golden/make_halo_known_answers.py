"""Known answers for the halo-exchange tables from the REFERENCE's own halo updater.

Runs `CubedSphereCommunicator.halo_update / vector_halo_update / synchronize_vector_interfaces` of the unmodified
reference (util/pace/util/communicator.py:331-555, halo_updater.py, halo_data_transformer.py, rotate.py) on
index-encoded fields — value = comp_code + rank*1e4 + i*100 + j with comp_code 1.0 for the x-field and 0.5 for the
y-field of a pair, so that source rank, source point, swapped component and sign can all be read off a received
value — for layouts 1, 2 and 3 (ranks as threads over oracle/refshim ThreadComm), and stores the resulting arrays
(level 0 only; all levels are equal) in tests/golden/topology/halo_known_answers.npz.

    PYTHONPATH=/root/repo python tests/golden/make_halo_known_answers.py       # build container only
"""
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NX, NZ, HALO = 4, 2, 3

CASES = {
    # name: (dims_x, dims_y or None, n_halo, mode)
    "scalar_cell_h3": (("x", "y", "z"), None, 3, "halo"),
    "scalar_cell_h2": (("x", "y", "z"), None, 2, "halo"),
    "scalar_cell_h1": (("x", "y", "z"), None, 1, "halo"),
    "scalar_corner_h3": (("x_interface", "y_interface", "z"), None, 3, "halo"),
    "scalar_zi_h3": (("x", "y", "z_interface"), None, 3, "halo"),
    "vector_dgrid_h3": (("x", "y_interface", "z"), ("x_interface", "y", "z"), 3, "halo"),
    "vector_cgrid_h3": (("x_interface", "y", "z"), ("x", "y_interface", "z"), 3, "halo"),
    "vector_agrid_h3": (("x", "y", "z"), ("x", "y", "z"), 3, "halo"),
    "vector_dgrid_h1": (("x", "y_interface", "z"), ("x_interface", "y", "z"), 1, "halo"),
    "sync_interfaces_dgrid": (("x", "y_interface", "z"), ("x_interface", "y", "z"), 0, "interface"),
}


def encode(shape, rank, code):
    i, j = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    return np.broadcast_to((code + rank * 1e4 + i * 100 + j)[:, :, None], shape).copy()


def run_layout(layout):
    from oracle.refshim import shim  # noqa: F401
    from oracle.refshim.threadcomm import ThreadComm, World
    import pace.util

    total = 6 * layout * layout
    world = World(total)
    results = {name: [None] * total for name in CASES}
    errors = []

    def work(rank):
        try:
            partitioner = pace.util.CubedSpherePartitioner(pace.util.TilePartitioner((layout, layout)))
            comm = pace.util.CubedSphereCommunicator(ThreadComm(world, rank), partitioner)
            sizer = pace.util.SubtileGridSizer.from_tile_params(
                nx_tile=NX * layout, ny_tile=NX * layout, nz=NZ, n_halo=HALO, extra_dim_lengths={}, layout=(layout, layout),
                tile_partitioner=partitioner.tile, tile_rank=comm.tile.rank)
            qf = pace.util.QuantityFactory.from_backend(sizer=sizer, backend="numpy")
            for name, (dx, dy, nh, mode) in CASES.items():
                qx = qf.zeros(list(dx), "m")
                qx.data[:] = encode(qx.data.shape, rank, 1.0)
                qy = None
                if dy is not None:
                    qy = qf.zeros(list(dy), "m")
                    qy.data[:] = encode(qy.data.shape, rank, 0.5)
                world.barrier.wait()
                if mode == "interface":
                    comm.synchronize_vector_interfaces(qx, qy)
                elif qy is None:
                    comm.halo_update(qx, nh)
                else:
                    comm.vector_halo_update(qx, qy, nh)
                world.barrier.wait()
                results[name][rank] = (np.array(qx.data[:, :, 0]), None if qy is None else np.array(qy.data[:, :, 0]))
        except BaseException as e:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            errors.append(e)
            world.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(total)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errors:
        raise errors[0]
    return results


if __name__ == "__main__":
    out = {}
    for layout in (1, 2, 3):
        res = run_layout(layout)
        for name, per_rank in res.items():
            out[f"L{layout}.{name}.x"] = np.stack([p[0] for p in per_rank])
            if per_rank[0][1] is not None:
                out[f"L{layout}.{name}.y"] = np.stack([p[1] for p in per_rank])
    dst = os.path.join(HERE, "topology", "halo_known_answers.npz")
    np.savez_compressed(dst, **out)
    print(len(out), "arrays ->", dst, os.path.getsize(dst) / 1e3, "kB")
