"""One process per GPU: the subdomains of the cube are split over world_size processes and halos cross processes by
pack -> torch.distributed send/recv -> unpack.  Run here with the gloo backend and 2 CPU processes on the host
simulation; the result must equal the single-process run bit for bit (the exchange is pure data movement)."""
import os
import sys
from datetime import timedelta

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rank, world, port, nx, layout, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    from oracle import hostsim

    hostsim.install()
    pc = None
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from pace_b200.util.communicator import ProcessComm

        pc = ProcessComm.from_torch_distributed()
    import bench

    dycore, state, comm, rt, gd = bench.build_dycore(nx, layout, 79, 1, 2, "cpu", pc, all_tracers=True)
    dycore.step_dynamics(state)
    out = {n: getattr(state, n).numpy() for n in ("u", "v", "w", "delp", "pt", "delz", "qvapor", "qsgs_tke", "ua")}
    np.savez(os.path.join(out_dir, f"w{world}_r{rank}.npz"), ranks=np.asarray(comm.local_ranks), **out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,layout", [(12, 1), (24, 2)])
def test_two_processes_equal_one_process(tmp_path, nx, layout):
    if torch.cuda.is_available():
        pytest.skip("CPU (gloo) test of the host-side exchange logic")
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_run, args=(1, port, nx, layout, str(tmp_path)), nprocs=1, join=True)
    mp.spawn(_run, args=(2, port + 1, nx, layout, str(tmp_path)), nprocs=2, join=True)
    one = np.load(tmp_path / "w1_r0.npz")
    for r in range(2):
        two = np.load(tmp_path / f"w2_r{r}.npz")
        ranks = two["ranks"]
        for k in two.files:
            if k == "ranks":
                continue
            np.testing.assert_array_equal(two[k], one[k][ranks], err_msg=f"{k} on process {r}")
