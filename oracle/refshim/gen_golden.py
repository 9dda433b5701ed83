"""Generate golden vectors from the UNMODIFIED reference (numpy backend) — run in the build container only.

    PYTHONPATH=/root/repo python -m oracle.refshim.gen_golden --nx 12 --layout 1 --out /tmp/pace_b200_golden/c12

Writes (np.savez, fp64, full arrays incl. halos, reference memory order [i, j, k]):
  grid_rank{r}.npz     every GridData / DampingCoefficients term of rank r
  state0_rank{r}.npz   analytic baroclinic initial DycoreState (after the init halo updates)
  state1_rank{r}.npz   DycoreState after `nsteps` calls of DynamicalCore.step_dynamics
  stage_rank{r}/<Stage>#<n>.npz   "in.<arg>" / "out.<arg>" snapshots around the n-th call of each stage
The committed subset under tests/golden/ is produced from this cache by tests/golden/make_committed.py.
"""
import argparse
import json
import os
import time

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=12)
    ap.add_argument("--layout", type=int, default=1)
    ap.add_argument("--nsteps", type=int, default=1)
    ap.add_argument("--capture-ranks", type=int, nargs="*", default=[0])
    ap.add_argument("--out", required=True)
    ap.add_argument("--n-split", type=int, default=1)
    ap.add_argument("--k-split", type=int, default=1)
    ap.add_argument("--capture-step", type=int, default=0, help="0-based index of the step whose stages are captured")
    ap.add_argument("--fill-tracers", action="store_true",
                    help="BASELINE configs[3]: tracer m (1..7) = qvapor * (m + 1) / 10 before the first step (SURVEY.md §8d)")
    ap.add_argument("--stages", nargs="*", default=None, help="capture only these stages (default: all)")
    ap.add_argument("--do-sat-adj", action="store_true", help="the stock baroclinic_c12.yaml setting (row f1)")
    ap.add_argument("--hord", type=int, default=None, help="override hord_dp = hord_tm = hord_vt = hord_mt (5: the other monotonic switch)")
    args = ap.parse_args()

    from oracle.refshim import runner

    os.makedirs(args.out, exist_ok=True)
    t0 = time.time()
    overrides = dict(n_split=args.n_split, k_split=args.k_split)
    if args.do_sat_adj:
        overrides["do_sat_adj"] = True
    if args.hord is not None:
        overrides.update(hord_dp=args.hord, hord_tm=args.hord, hord_vt=args.hord, hord_mt=args.hord)
    ctxs, cap = runner.run(
        args.nx, (args.layout, args.layout), nsteps=args.nsteps, capture_ranks=tuple(args.capture_ranks),
        config_overrides=overrides, capture_step=args.capture_step, stages=args.stages,
        on_built=runner.fill_tracers if args.fill_tracers else None,
    )
    for ctx in ctxs:
        r = ctx["rank"]
        np.savez(os.path.join(args.out, f"grid_rank{r}.npz"), **runner.grid_arrays(ctx))
        np.savez(os.path.join(args.out, f"state0_rank{r}.npz"), **ctx.get("state_before_capture", ctx["state0"]))
        np.savez(os.path.join(args.out, f"state1_rank{r}.npz"), **runner.state_arrays(ctx["state"]))
    for r, stages in cap.data.items():
        d = os.path.join(args.out, f"stage_rank{r}")
        os.makedirs(d, exist_ok=True)
        for key, rec in stages.items():
            flat = {f"in.{k}": v for k, v in rec["in"].items()}
            flat.update({f"out.{k}": v for k, v in rec["out"].items()})
            np.savez(os.path.join(d, key + ".npz"), **flat)
    meta = dict(capture_step=args.capture_step, nx=args.nx, layout=args.layout, nsteps=args.nsteps, n_split=args.n_split, k_split=args.k_split,
                fill_tracers=bool(args.fill_tracers), do_sat_adj=bool(args.do_sat_adj), hord=args.hord, timing=ctxs[0].get("timing"), wall=time.time() - t0, config=runner.C12_CONFIG,
                reference="ai2cm/pace @ /root/reference, numpy backend via oracle/refshim")
    with open(os.path.join(args.out, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, default=str)
    print("wrote", args.out, "in", time.time() - t0, "s")


if __name__ == "__main__":
    main()
