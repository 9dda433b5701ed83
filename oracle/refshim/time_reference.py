"""Time the ACTUAL reference (ai2cm/pace, vendored GT4Py numpy backend, ranks as threads over ThreadComm) on the host
cores of the build container — the reference cannot travel to the GPU box, so this number is recorded once per round in
profiles/reference_numpy_timing.json and carried by bench.py as `cpu_baseline.reference_numpy`.

    PYTHONPATH=/root/repo python -m oracle.refshim.time_reference [--nx 24 --layout 2 --k-split 2 --n-split 6 --steps 1]

Workload: the benchmark's decomposition and split at a size the numpy backend finishes in minutes (c24 L79, layout (2,2) =
24 ranks, k_split=2, n_split=6, 8 non-zero tracers, do_sat_adj off); one warm-up step (GT4Py stencil build + caches), then
`--steps` timed `DynamicalCore.step_dynamics`.  The per-cell cost and its extrapolation to C128 by cell count are written
alongside (the numpy backend's cost per cell falls with subdomain size, so the extrapolation is an upper bound).
"""
import argparse
import json
import os
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=24)
    ap.add_argument("--layout", type=int, default=2)
    ap.add_argument("--k-split", type=int, default=2)
    ap.add_argument("--n-split", type=int, default=6)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                                                  "profiles", "reference_numpy_timing.json"))
    args = ap.parse_args()
    from oracle.refshim import runner

    t0 = time.time()
    ctxs, _ = runner.run(args.nx, (args.layout, args.layout), nsteps=1 + args.steps, capture_ranks=(), stages=["NONE"],
                         config_overrides=dict(n_split=args.n_split, k_split=args.k_split), on_built=runner.fill_tracers, verbose=True)
    timing = ctxs[0]["timing"]
    per_step = sum(timing[1:]) / len(timing[1:])
    cells = 6 * args.nx * args.nx * 79
    out = {
        "what": "unmodified ai2cm/pace DynamicalCore.step_dynamics, GT4Py numpy backend, ranks as threads (oracle/refshim)",
        "workload": f"c{args.nx} L79 layout ({args.layout},{args.layout}), k_split={args.k_split}, n_split={args.n_split}, 8 non-zero tracers, do_sat_adj off",
        "host_cores": os.cpu_count(), "threads": 6 * args.layout ** 2, "warmup_steps": 1, "timed_steps": len(timing[1:]),
        "seconds_per_step": per_step, "first_step_seconds_incl_stencil_build": timing[0],
        "microseconds_per_cell_step": per_step / cells * 1e6,
        "extrapolated_C128_seconds_per_step_by_cell_count": per_step * (128 / args.nx) ** 2,
        "wall_seconds_total": time.time() - t0, "where": "build container (no GPU); python threads share the GIL outside numpy kernels",
    }
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
