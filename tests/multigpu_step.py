"""One strict step case on SEVERAL GPUs (launched by torch.distributed.run, one process per GPU): the inter-GPU halo path
(fv3_halo_pack_segments -> torch.distributed send/recv, or fv3_halo_exchange_nccl with FV3_NATIVE_NCCL=1 ->
fv3_halo_unpack_segments on the communication stream) against the
reference's final state.  Every process checks the reference ranks it owns; exit code 0 = all within tolerance.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tests/multigpu_step.py c24L2k2n3
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    case = sys.argv[1] if len(sys.argv) > 1 else "c24L2k2n3"
    use_cuda = torch.cuda.is_available() and os.environ.get("FV3_MULTI_CPU") != "1"
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if use_cuda:
        torch.cuda.set_device(local_rank)
        dev = f"cuda:{local_rank}"
        dist.init_process_group("nccl", device_id=torch.device(dev))
    else:
        from oracle import hostsim

        hostsim.install(openmp=False)
        dev = "cpu"
        dist.init_process_group("gloo")
    from pace_b200.util.communicator import ProcessComm
    from tests import step_cases as S

    pc = ProcessComm.from_torch_distributed()
    meta, grids, s0, s1 = S.load(case)
    dycore, state = S.build(meta, grids, s0, dev, process_comm=pc)
    comm = dycore.comm
    dycore.step_dynamics(state)
    if use_cuda:
        torch.cuda.synchronize()
    out = state.as_numpy()
    mine = {r: z for r, z in s1.items() if r in comm.local_ranks}
    failures, achieved = S.compare(out, mine, meta, first_rank=comm.first_rank)
    worst = max([v[0] for v in achieved.values()] + [0.0])
    transport = "fv3_halo_exchange_nccl" if pc.nccl_comm is not None else "torch.distributed"
    want = os.environ.get("FV3_EXPECT_TRANSPORT")
    if want and want != transport:
        failures.append(f"halo messages went through {transport}, expected {want}")
    print(f"[rank {pc.rank}] {case}: reference ranks checked {sorted(mine)}, worst relative error {worst:.2e}, "
          f"halo messages via {transport}, {len(failures)} failures", flush=True)
    for f in failures[:10]:
        print(f"[rank {pc.rank}]   {f}", flush=True)
    bad = torch.tensor([len(failures)], dtype=torch.int64, device=dev)
    dist.all_reduce(bad)
    sys.stdout.flush()
    if use_cuda:
        torch.cuda.synchronize()
    os._exit(0 if int(bad.item()) == 0 else 1)


if __name__ == "__main__":
    main()
