"""Reduce full reference dumps (oracle/refshim/gen_golden.py, numpy backend of the unmodified reference) to the
committed STRICT step cases: the reference's own grid + initial state go in, its final state comes out, no tolerance
for generator differences is needed.

    # BASELINE configs[2]/[3] split (k_split=2, n_split=6, 8 non-zero tracers) at c12 layout (1,1):
    PYTHONPATH=/root/repo python -m oracle.refshim.gen_golden --nx 12 --layout 1 --k-split 2 --n-split 6 --fill-tracers \
        --capture-ranks 0 --stages Tracer2D1L Remapping Riem_Solver3 FVSetup D_SW --out $CACHE/c12k2n6
    # the benchmark's decomposition, layout (2,2), multi-substep, 8 non-zero tracers:
    PYTHONPATH=/root/repo python -m oracle.refshim.gen_golden --nx 24 --layout 2 --k-split 2 --n-split 3 --fill-tracers \
        --capture-ranks --stages NONE --out $CACHE/c24L2k2n3
    python tests/golden/make_step_strict.py

Written:
  tests/golden/c12k2n6_step/     meta + state1 of ranks 0 and 3 (all levels, prognostic fields + 8 tracers); inputs are
                                 tests/golden/c12_step (identical grid and state0; tracers 1..7 follow the fill rule)
  tests/golden/c24L2_inputs/     grid + state0 of all 24 ranks of the c24 layout (2,2) case (non-zero fields only)
  tests/golden/c24L2k2n3_step/   meta + state1 of ranks 0, 5, 14, 23 on a level subsample
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.environ.get("PACE_B200_GOLDEN_CACHE", "/tmp/pace_b200_golden")
TRACERS = ["qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke"]
FIELDS = ["u", "v", "w", "delz", "delp", "pt", "pe", "peln", "pk", "pkz", "ps", "q_con", "omga", "ua", "va"] + TRACERS
LEVELS = sorted(set(list(range(0, 80, 4)) + [77, 78, 79]))


def size(d):
    return sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)) / 1e6


def c12k2n6():
    src, dst, inp = os.path.join(CACHE, "c12k2n6"), os.path.join(HERE, "c12k2n6_step"), os.path.join(HERE, "c12_step")
    os.makedirs(dst, exist_ok=True)
    for r in range(6):  # the committed c12_step inputs ARE this run's inputs (tracers 1..7 aside)
        a, b = np.load(os.path.join(inp, f"state0_rank{r}.npz")), np.load(os.path.join(src, f"state0_rank{r}.npz"))
        for k in b.files:
            if k in TRACERS[1:]:
                m = TRACERS.index(k)
                assert np.array_equal(b[k], b["qvapor"] * (m + 1) * 0.1), k
            else:
                assert np.array_equal(a[k], b[k], equal_nan=True), (r, k)
        ga, gb = np.load(os.path.join(inp, f"grid_rank{r}.npz")), np.load(os.path.join(src, f"grid_rank{r}.npz"))
        assert all(np.array_equal(ga[k], gb[k], equal_nan=True) for k in gb.files)
    meta = json.load(open(os.path.join(src, "meta.json")))
    meta.update(inputs="c12_step", fields=FIELDS, ranks=[0, 3], levels=None)
    json.dump(meta, open(os.path.join(dst, "meta.json"), "w"), indent=1, default=str)
    for r in (0, 3):
        z = np.load(os.path.join(src, f"state1_rank{r}.npz"))
        np.savez_compressed(os.path.join(dst, f"state1_rank{r}.npz"), **{n: z[n] for n in FIELDS})
    print(dst, f"{size(dst):.1f} MB")


# savepoint variable -> key in the reference's stage capture (oracle/refshim/runner.py wraps each stage's __call__)
CHECKPOINTS = {
    "Tracer2D1L-In": ("Tracer2D1L", "in", dict(dp1="dp1", mfxd="x_mass_flux", mfyd="y_mass_flux", cxd="x_courant", cyd="y_courant")),
    "Tracer2D1L-Out": ("Tracer2D1L", "out", dict(dp1="dp1", mfxd="x_mass_flux", mfyd="y_mass_flux", cxd="x_courant", cyd="y_courant",
                                                 **{t: "tracers." + t for t in TRACERS})),
    "Remapping-In": ("Remapping", "in", {n: n for n in ("pt", "delp", "delz", "peln", "u", "v", "w", "cappa", "pk", "pe", "ps", "wsd", "dp1")}),
    "Remapping-Out": ("Remapping", "out", {n: n for n in ("pt", "delp", "delz", "peln", "u", "v", "w", "cappa", "pkz", "pk", "pe", "dp1")}),
    # ucd, vcd and divgdd are left out: the reference's d_sw uses uc, vc and divgd as work arrays, what they hold on
    # exit is scratch that nothing reads (c_sw rewrites all three)
    "D_SW-Out": ("D_SW", "out", dict(wd="w", delpd="delp", ud="u", vd="v", ptd="pt", uad="ua", vad="va",
                                     xfxd="xfx", yfxd="yfx", mfxd="mfx", mfyd="mfy")),
}
CHECKPOINT_CALLS = {"Tracer2D1L": (0, 1), "Remapping": (0, 1), "D_SW": (0, 7, 11)}


def c12k2n6_checkpoints():
    """In-flight savepoints of the k_split=2, n_split=6 run, rank 0: the arguments the reference's TracerAdvection,
    LagrangianToEulerian and DGridShallowWaterLagrangianDynamics calls entered / left with, cropped to the compute
    domain (+1 for staggered fields) on a level subsample -> tests/golden/c12k2n6_step/checkpoints_rank0.npz, keyed
    "<savepoint>#<call>/<variable>" with the savepoint and variable names of the reference's checkpointer calls
    (fv_dynamics.py:340-422, dyn_core.py:608-668)."""
    src, dst = os.path.join(CACHE, "c12k2n6", "stage_rank0"), os.path.join(HERE, "c12k2n6_step")
    out = {}
    for sp, (stage, side, names) in CHECKPOINTS.items():
        for call in CHECKPOINT_CALLS[stage]:
            z = np.load(os.path.join(src, f"{stage}#{call}.npz"))
            for var, key in names.items():
                a = z[f"{side}.{key}"]
                a = a[3:16, 3:16]
                if a.ndim == 3:
                    a = a[:, :, [k for k in LEVELS if k < (80 if key in ("pe", "peln", "pk") else 79)]]
                out[f"{sp}#{call}/{var}"] = np.ascontiguousarray(a)
    np.savez_compressed(os.path.join(dst, "checkpoints_rank0.npz"), levels=np.array(LEVELS), **out)
    print(os.path.join(dst, "checkpoints_rank0.npz"), f"{os.path.getsize(os.path.join(dst, 'checkpoints_rank0.npz')) / 1e6:.1f} MB")


def c24L2():
    src = os.path.join(CACHE, "c24L2k2n3")
    inp, dst = os.path.join(HERE, "c24L2_inputs"), os.path.join(HERE, "c24L2k2n3_step")
    os.makedirs(inp, exist_ok=True)
    os.makedirs(dst, exist_ok=True)
    old = os.path.join(CACHE, "c24L2")  # the k_split = n_split = 1 dump behind tests/golden/c24L2_step: same inputs
    for r in range(24):
        g = np.load(os.path.join(src, f"grid_rank{r}.npz"))
        np.savez_compressed(os.path.join(inp, f"grid_rank{r}.npz"), **{k: g[k] for k in g.files})
        z = np.load(os.path.join(src, f"state0_rank{r}.npz"))
        keep = {}
        for k in z.files:
            if k in TRACERS[1:]:
                m = TRACERS.index(k)
                assert np.array_equal(z[k], z["qvapor"] * (m + 1) * 0.1), k
                continue
            if np.any(z[k] != 0):
                keep[k] = z[k]
        np.savez_compressed(os.path.join(inp, f"state0_rank{r}.npz"), **keep)
        if os.path.exists(os.path.join(old, f"state0_rank{r}.npz")):
            zo = np.load(os.path.join(old, f"state0_rank{r}.npz"))
            for k in keep:
                assert np.array_equal(zo[k], keep[k], equal_nan=True), ("c24L2 k1n1 dump has different inputs", r, k)
    json.dump(dict(nx=24, layout=2, zero_fields="fields absent from state0_rank*.npz are zero; tracers 1..7 follow the fill rule "
                   "when a case says fill_tracers", reference="ai2cm/pace numpy backend via oracle/refshim"),
              open(os.path.join(inp, "meta.json"), "w"), indent=1)
    meta = json.load(open(os.path.join(src, "meta.json")))
    ranks = [0, 5, 14, 23]
    meta.update(inputs="c24L2_inputs", fields=FIELDS, ranks=ranks, levels=LEVELS)
    json.dump(meta, open(os.path.join(dst, "meta.json"), "w"), indent=1, default=str)
    for r in ranks:
        z = np.load(os.path.join(src, f"state1_rank{r}.npz"))
        np.savez_compressed(os.path.join(dst, f"state1_rank{r}.npz"),
                            **{n: (z[n][:, :, LEVELS] if z[n].ndim == 3 else z[n]) for n in FIELDS})
    print(inp, f"{size(inp):.1f} MB;", dst, f"{size(dst):.1f} MB")


def c12_sat():
    """do_sat_adj=True (row f1): the stock baroclinic_c12.yaml (`c12sat`, k_split = n_split = 1, qvapor only) and a
    k_split = 2 run with 8 non-zero tracers (`c12satk2`: both last_step branches, every phase-change branch that
    condensate can take), plus the SatAdjust3d call snapshots of rank 0 for the stage test."""
    import sys

    sys.path.insert(0, HERE)
    from make_committed import reduce_stage

    inp = os.path.join(HERE, "c12_step")
    for case in ("c12sat", "c12satk2", "c12hord5"):
        src, dst = os.path.join(CACHE, case), os.path.join(HERE, case + "_step")
        if not os.path.exists(os.path.join(src, "meta.json")):
            continue
        os.makedirs(dst, exist_ok=True)
        meta = json.load(open(os.path.join(src, "meta.json")))
        for r in range(6):
            a, b = np.load(os.path.join(inp, f"state0_rank{r}.npz")), np.load(os.path.join(src, f"state0_rank{r}.npz"))
            for k in b.files:
                if meta.get("fill_tracers") and k in TRACERS[1:]:
                    continue
                assert np.array_equal(a[k], b[k], equal_nan=True), (case, r, k)
        meta.update(inputs="c12_step", fields=FIELDS + ["qcld"], ranks=[0, 3], levels=None)
        json.dump(meta, open(os.path.join(dst, "meta.json"), "w"), indent=1, default=str)
        for r in (0, 3):
            z = np.load(os.path.join(src, f"state1_rank{r}.npz"))
            np.savez_compressed(os.path.join(dst, f"state1_rank{r}.npz"), **{n: z[n] for n in FIELDS + ["qcld"]})
        sd = os.path.join(dst, "stage_rank0")
        os.makedirs(sd, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(src, "stage_rank0"))) if os.path.isdir(os.path.join(src, "stage_rank0")) else []:
            if f.startswith("SatAdjust3d"):
                reduce_stage(os.path.join(src, "stage_rank0", f), os.path.join(sd, f))
        print(dst, f"{sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(dst) for f in fs) / 1e6:.1f} MB")


if __name__ == "__main__":
    c12_sat()
    if os.path.exists(os.path.join(CACHE, "c12k2n6", "meta.json")):
        c12k2n6()
        c12k2n6_checkpoints()
    if os.path.exists(os.path.join(CACHE, "c24L2k2n3", "meta.json")):
        c24L2()
