"""Golden vectors of AdjustNegativeTracerMixingRatio (neg_adj3.py:316-420) from the UNMODIFIED reference (numpy
backend) — run in the build container only:

    PYTHONPATH=/root/repo python -m oracle.refshim.gen_neg_adj --out tests/golden/neg_adj3

The baroclinic test case never produces negative mixing ratios, so the inputs are synthetic: seeded random fields of
realistic magnitude in which a fraction of the values of every species is negative (case 0) and a case whose
columns are all non-negative (case 1: the call must leave every field bit-identical).  Stored: compute-domain arrays
[i, j, k] of the 9 arguments before ("in.") and after ("out.") the call.
"""
import argparse
import os

import numpy as np

from . import shim  # noqa: F401

import pace.dsl.stencil  # noqa: E402
import pace.util  # noqa: E402
from pace.dsl.dace.dace_config import DaceConfig  # noqa: E402
from pace.fv3core.stencils.neg_adj3 import AdjustNegativeTracerMixingRatio  # noqa: E402
from pace.util import X_DIM, Y_DIM, Z_DIM  # noqa: E402

NAMES = ["qvapor", "qliquid", "qrain", "qsnow", "qice", "qgraupel", "qcld", "pt", "delp"]


def make_inputs(rng, nx, ny, nz, neg_fraction):
    shape = (nx, ny, nz)
    out = {}
    scale = dict(qvapor=3e-3, qliquid=2e-5, qrain=1e-5, qsnow=1e-5, qice=2e-5, qgraupel=1e-5, qcld=0.3)
    for n, s in scale.items():
        v = s * rng.random(shape) ** 2
        if neg_fraction > 0:
            neg = rng.random(shape) < neg_fraction
            v = np.where(neg, -0.5 * s * rng.random(shape), v)
            zero = rng.random(shape) < 0.1
            v = np.where(zero, 0.0, v)
        out[n] = v
    if neg_fraction > 0:
        # whole columns in deficit, negative top / bottom levels
        out["qvapor"][0, 0, :] = -1e-4
        out["qvapor"][1, :, 0] = -2e-4
        out["qvapor"][2, :, -1] = -3e-4
        out["qrain"][3, 0, :] = -1e-6
        out["qcld"][1, 1, -1] = -0.2
        out["qcld"][2, 2, 0] = -0.1
    out["pt"] = 210.0 + 90.0 * rng.random(shape)
    out["delp"] = 50.0 + 1500.0 * rng.random(shape)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--nx", type=int, default=12)
    ap.add_argument("--nz", type=int, default=79)
    args = ap.parse_args()
    backend = "numpy"
    nx = ny = args.nx
    dace_config = DaceConfig(communicator=None, backend=backend)
    stencil_config = pace.dsl.stencil.StencilConfig(
        compilation_config=pace.dsl.stencil.CompilationConfig(backend=backend, rebuild=False, validate_args=False),
        dace_config=dace_config,
    )
    sizer = pace.util.SubtileGridSizer(nx=nx, ny=ny, nz=args.nz, n_halo=3, extra_dim_lengths={})
    gi = pace.dsl.stencil.GridIndexing(domain=(nx, ny, args.nz), n_halo=3, south_edge=True, north_edge=True,
                                       west_edge=True, east_edge=True)
    qf = pace.util.QuantityFactory.from_backend(sizer=sizer, backend=backend)
    sf = pace.dsl.stencil.StencilFactory(config=stencil_config, grid_indexing=gi)
    obj = AdjustNegativeTracerMixingRatio(sf, quantity_factory=qf, check_negative=False, hydrostatic=False)
    os.makedirs(args.out, exist_ok=True)
    for case, (seed, frac) in enumerate([(11, 0.3), (13, 0.0)]):
        rng = np.random.default_rng(seed)
        inp = make_inputs(rng, nx, ny, args.nz, frac)
        qs = {}
        for n in NAMES:
            q = qf.zeros([X_DIM, Y_DIM, Z_DIM], units="unknown")
            q.view[:] = inp[n]
            qs[n] = q
        obj(*[qs[n] for n in NAMES])
        rec = {"in." + n: inp[n] for n in NAMES}
        rec.update({"out." + n: np.array(qs[n].view[:], copy=True) for n in NAMES})
        np.savez_compressed(os.path.join(args.out, f"case{case}.npz"), **rec)
        changed = {n: int((rec["out." + n] != rec["in." + n]).sum()) for n in NAMES}
        print("case", case, "changed values:", changed, "min out:", {n: float(rec["out." + n].min()) for n in NAMES[:7]})


if __name__ == "__main__":
    main()
