"""Dump the reference's metric terms and analytic baroclinic initial state (no dycore build) — build container only.

    PYTHONPATH=/root/repo python -m oracle.refshim.gen_grid --nx 24 --layout 2 --out /tmp/pace_b200_golden/grid_c24L2

Used to validate pace_b200.util.grid.generation and pace_b200.fv3core.initialization.baroclinic
(tests/test_grid_generation.py); a reduced copy is committed under tests/golden/grid_*.
"""
import argparse
import os

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=12)
    ap.add_argument("--layout", type=int, default=1)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    from oracle.refshim import runner

    os.makedirs(args.out, exist_ok=True)
    ctxs, _ = runner.run(args.nx, (args.layout, args.layout), build_only=True, capture_ranks=())
    for ctx in ctxs:
        r = ctx["rank"]
        np.savez(os.path.join(args.out, f"grid_rank{r}.npz"), **runner.grid_arrays(ctx))
        np.savez(os.path.join(args.out, f"state0_rank{r}.npz"), **ctx["state0"])
    print("wrote", args.out)


if __name__ == "__main__":
    main()
