#!/usr/bin/env python
"""One `step_dynamics` of the bench workload bracketed by cudaProfilerStart/Stop, for use under ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--nx 128 --k-split 1 --n-split 1]

Without ncu it just runs the step (and prints the eager device time).  Profiling aid only, not a bench.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=128)
    ap.add_argument("--layout", type=int, default=2)
    ap.add_argument("--k-split", type=int, default=1)
    ap.add_argument("--n-split", type=int, default=1)
    ap.add_argument("--warm", type=int, default=1)
    args = ap.parse_args()
    import torch

    import bench

    dycore, state, comm, rt, gd = bench.build_dycore(args.nx, args.layout, 79, args.k_split, args.n_split, "cuda:0")
    for _ in range(args.warm):
        dycore.step_dynamics(state)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    e0.record()
    dycore.step_dynamics(state)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"step: {e0.elapsed_time(e1):.3f} ms", file=sys.stderr)


if __name__ == "__main__":
    main()
