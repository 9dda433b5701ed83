"""Reduce the local reference dump (tests/golden/_cache/<case>, made by oracle/refshim/gen_golden.py) to the
committed golden subset tests/golden/<case>/.

Kept per stage: every array/scalar input, and every output that differs from its input (unchanged outputs are
dropped; the loader falls back to the input).  Files are np.savez_compressed, fp64, reference order [i, j, k].

    python tests/golden/make_committed.py c12 0 Riem_Solver_C#0 C_SW#0 ...
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def reduce_stage(src, dst):
    z = np.load(src)
    out = {}
    for k in z.files:
        if k.startswith("in."):
            out[k] = z[k]
    for k in z.files:
        if k.startswith("out."):
            kin = "in." + k[4:]
            if kin in z.files and z[kin].shape == z[k].shape and np.array_equal(z[kin], z[k], equal_nan=True):
                continue
            out[k] = z[k]
    np.savez_compressed(dst, **out)


def main():
    case, rank = sys.argv[1], int(sys.argv[2])
    stages = sys.argv[3:]
    src = os.path.join(os.environ.get("PACE_B200_GOLDEN_CACHE", "/tmp/pace_b200_golden"), case)
    dst = os.path.join(HERE, case)
    os.makedirs(os.path.join(dst, f"stage_rank{rank}"), exist_ok=True)
    shutil.copy(os.path.join(src, "meta.json"), os.path.join(dst, "meta.json"))
    g = np.load(os.path.join(src, f"grid_rank{rank}.npz"))
    np.savez_compressed(os.path.join(dst, f"grid_rank{rank}.npz"), **{k: g[k] for k in g.files})
    for st in stages:
        reduce_stage(os.path.join(src, f"stage_rank{rank}", st + ".npz"), os.path.join(dst, f"stage_rank{rank}", st + ".npz"))
    total = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(dst) for f in fs)
    print(f"{dst}: {total / 1e6:.1f} MB")


if __name__ == "__main__":
    main()
