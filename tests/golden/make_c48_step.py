"""Reduce the reference's C48 one-step dump (oracle/refshim/gen_golden.py --nx 48 --layout 1 --capture-ranks, written to
$PACE_B200_GOLDEN_CACHE/c48) to the committed subset tests/golden/c48_step/: for ranks 0 and 3 the prognostic fields
before (state0) and after (state1) one DynamicalCore.step_dynamics, on a subsample of levels (every 6th and the last
two), fp64, reference order [i, j, k] with halos.

    python tests/golden/make_c48_step.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.environ.get("PACE_B200_GOLDEN_CACHE", "/tmp/pace_b200_golden")
FIELDS = ["u", "v", "w", "delz", "delp", "pt", "qvapor", "ua", "va", "omga"]
LEVELS = sorted(set(list(range(0, 79, 6)) + [77, 78]))


def main():
    src, dst = os.path.join(CACHE, "c48"), os.path.join(HERE, "c48_step")
    os.makedirs(dst, exist_ok=True)
    meta = json.load(open(os.path.join(src, "meta.json")))
    meta["levels"] = LEVELS
    meta["fields"] = FIELDS
    json.dump(meta, open(os.path.join(dst, "meta.json"), "w"), indent=1, default=str)
    for r in (0, 3):
        for which in ("state0", "state1"):
            z = np.load(os.path.join(src, f"{which}_rank{r}.npz"))
            out = {}
            for n in FIELDS:
                a = z[n]
                out[n] = a[:, :, LEVELS] if a.ndim == 3 else a
            if which == "state0":
                out["phis"] = z["phis"]
            np.savez_compressed(os.path.join(dst, f"{which}_rank{r}.npz"), **out)
    print("wrote", dst, sum(os.path.getsize(os.path.join(dst, f)) for f in os.listdir(dst)) / 1e6, "MB")


if __name__ == "__main__":
    main()
