"""Reduce the reference's c24 layout (2,2) one-step dump (oracle/refshim/gen_golden.py --nx 24 --layout 2
--capture-ranks, written to $PACE_B200_GOLDEN_CACHE/c24L2) to the committed subset tests/golden/c24L2_step/: for four
ranks on four different tiles (one per position inside a tile: 0 = tile 0 SW, 5 = tile 1 SE, 14 = tile 3 NW,
23 = tile 5 NE) the prognostic fields before and after one DynamicalCore.step_dynamics on a subsample of levels.

    python tests/golden/make_layout2_step.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.environ.get("PACE_B200_GOLDEN_CACHE", "/tmp/pace_b200_golden")
FIELDS = ["u", "v", "w", "delz", "delp", "pt", "qvapor", "ua", "va", "omga"]
LEVELS = sorted(set(list(range(0, 79, 4)) + [77, 78]))
RANKS = (0, 5, 14, 23)


def main():
    src, dst = os.path.join(CACHE, "c24L2"), os.path.join(HERE, "c24L2_step")
    os.makedirs(dst, exist_ok=True)
    meta = json.load(open(os.path.join(src, "meta.json")))
    meta["levels"] = LEVELS
    meta["fields"] = FIELDS
    meta["ranks"] = list(RANKS)
    json.dump(meta, open(os.path.join(dst, "meta.json"), "w"), indent=1, default=str)
    for r in RANKS:
        for which in ("state0", "state1"):
            z = np.load(os.path.join(src, f"{which}_rank{r}.npz"))
            out = {n: (z[n][:, :, LEVELS] if z[n].ndim == 3 else z[n]) for n in FIELDS}
            np.savez_compressed(os.path.join(dst, f"{which}_rank{r}.npz"), **out)
    print("wrote", dst, sum(os.path.getsize(os.path.join(dst, f)) for f in os.listdir(dst)) / 1e6, "MB")


if __name__ == "__main__":
    main()
