"""Run the UNMODIFIED reference dycore (numpy backend) for all ranks in one process and capture stage I/O.

TEST INFRASTRUCTURE ONLY (see oracle/refshim/__init__.py).  Ranks run as Python threads over
`ThreadComm`.  Every stage listed in STAGE_CLASSES (reference class `__call__`) and in
ACOUSTIC_STENCILS (FrozenStencil attributes of `AcousticDynamics`) is wrapped so that all array
arguments are snapshotted before ("In") and after ("Out") the call.

Construction follows /root/reference/tests/main/fv3core/test_dycore_call.py:22-137.
"""
import inspect
import os
import threading
import time
from datetime import timedelta

import numpy as np

from . import shim  # noqa: F401
from .threadcomm import ThreadComm, World

import pace.dsl.stencil  # noqa: E402
import pace.fv3core.initialization.baroclinic as baroclinic_init  # noqa: E402
import pace.util  # noqa: E402
from pace import fv3core  # noqa: E402
from pace.dsl.dace.dace_config import DaceConfig  # noqa: E402
from pace.util.grid import DampingCoefficients, GridData, MetricTerms  # noqa: E402

_tls = threading.local()
_build_lock = threading.Lock()

# dycore_config of driver/examples/configs/baroclinic_c12.yaml:41-88 (do_sat_adj off for kernel parity)
C12_CONFIG = dict(
    ntiles=6, nwat=6, dt_atmos=225, a_imp=1.0, beta=0.0, consv_te=False, d2_bg=0.0, d2_bg_k1=0.2,
    d2_bg_k2=0.1, d4_bg=0.15, d_con=1.0, d_ext=0.0, dddmp=0.5, delt_max=0.002, do_sat_adj=False,
    do_vort_damp=True, fill=True, hord_dp=6, hord_mt=6, hord_tm=6, hord_tr=8, hord_vt=6,
    hydrostatic=False, k_split=1, ke_bg=0.0, kord_mt=9, kord_tm=-9, kord_tr=9, kord_wz=9, n_split=1,
    nord=3, p_fac=0.05, rf_fast=True, rf_cutoff=3000.0, tau=10.0, vtdm4=0.06, z_tracer=True, do_qa=True,
    n_sponge=48,
)


def make_config(nx, layout, npz=79, **overrides):
    kw = dict(C12_CONFIG)
    kw.update(overrides)
    return fv3core.DynamicalCoreConfig(layout=tuple(layout), npx=nx + 1, npy=nx + 1, npz=npz, **kw)


class Capture:
    """Thread-safe store of stage snapshots: captures[rank][f"{stage}#{n}"] = {"in": {...}, "out": {...}}."""

    def __init__(self, ranks, stages=None):
        self.ranks = set(ranks)
        self.stages = None if stages is None else set(stages)
        self.data = {r: {} for r in ranks}
        self.counts = {r: {} for r in ranks}
        self.enabled = True

    def want(self, stage):
        r = getattr(_tls, "rank", None)
        return self.enabled and r in self.ranks and (self.stages is None or stage in self.stages)


def _snap(v):
    if isinstance(v, pace.util.Quantity):
        return np.array(v.data, copy=True)
    if isinstance(v, np.ndarray):
        return np.array(v, copy=True)
    if isinstance(v, (bool, int, float, np.floating, np.integer)):
        return np.asarray(v)
    if isinstance(v, dict) and v and all(isinstance(x, pace.util.Quantity) for x in v.values()):
        return {k: _snap(x) for k, x in v.items()}
    if hasattr(v, "__dataclass_fields__"):
        out = {}
        for name in v.__dataclass_fields__:
            x = getattr(v, name)
            if isinstance(x, pace.util.Quantity):
                out[name] = _snap(x)
        return out or None
    return None


def _flatten(prefix, val, out):
    if val is None:
        return
    if isinstance(val, dict):
        for k, x in val.items():
            _flatten(f"{prefix}.{k}", x, out)
    else:
        out[prefix] = val


def wrap_callable(fn, stage, capture, skip_self=True, extra=None):
    """Return a wrapper of `fn` that snapshots array args before/after.  `extra(self)` -> dict of more arrays."""
    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        sig = None

    def wrapper(*args, **kwargs):
        if not capture.want(stage):
            return fn(*args, **kwargs)
        rank = _tls.rank
        n = capture.counts[rank].get(stage, 0)
        capture.counts[rank][stage] = n + 1
        names = {}
        if sig is not None:
            try:
                bound = sig.bind(*args, **kwargs)
                names = dict(bound.arguments)
            except TypeError:
                names = {f"arg{i}": a for i, a in enumerate(args)}
                names.update(kwargs)
        else:
            names = {f"arg{i}": a for i, a in enumerate(args)}
            names.update(kwargs)
        selfobj = names.pop("self", None)
        # FrozenStencil.__call__(*args, **kwargs): name positional args after the stencil definition's parameters
        if "args" in names and isinstance(names["args"], tuple):
            pos = names.pop("args")
            kw = names.pop("kwargs", {}) or {}
            argn = getattr(fn, "_argument_names", None) or getattr(getattr(fn, "__self__", None), "_argument_names", None) or [f"arg{i}" for i in range(len(pos))]
            names = dict(zip(argn, pos))
            names.update(kw)
        rec = {"in": {}, "out": {}}
        for k, v in names.items():
            _flatten(k, _snap(v), rec["in"])
        if extra is not None and selfobj is not None:
            for k, v in extra(selfobj).items():
                _flatten(k, _snap(v), rec["in"])
        ret = fn(*args, **kwargs)
        for k, v in names.items():
            _flatten(k, _snap(v), rec["out"])
        if extra is not None and selfobj is not None:
            for k, v in extra(selfobj).items():
                _flatten(k, _snap(v), rec["out"])
        if ret is not None:
            if isinstance(ret, tuple):
                for i, v in enumerate(ret):
                    _flatten(f"ret{i}", _snap(v), rec["out"])
            else:
                _flatten("ret", _snap(ret), rec["out"])
        capture.data[rank][f"{stage}#{n}"] = rec
        return ret

    wrapper.__wrapped__ = fn
    return wrapper


def _stage_classes():
    from pace.fv3core.stencils import (
        a2b_ord4, c_sw, d2a2c_vect, d_sw, del2cubed, delnflux, divergence_damping, dyn_core, fillz, fvtp2d,
        fxadv, map_single, mapn_tracer, neg_adj3, nh_p_grad, pk3_halo, ray_fast, remap_profile, remapping,
        riem_solver3, riem_solver_c, saturation_adjustment, tracer_2d_1l, updatedzc, updatedzd, xppm, xtp_u, yppm, ytp_v,
    )
    from pace.stencils import c2l_ord

    return {
        "AcousticDynamics": dyn_core.AcousticDynamics,
        "C_SW": c_sw.CGridShallowWaterDynamics,
        "D2A2C": d2a2c_vect.DGrid2AGrid2CGridVectors,
        "UpdateDzC": updatedzc.UpdateGeopotentialHeightOnCGrid,
        "Riem_Solver_C": riem_solver_c.NonhydrostaticVerticalSolverCGrid,
        "D_SW": d_sw.DGridShallowWaterLagrangianDynamics,
        "FxAdv": fxadv.FiniteVolumeFluxPrep,
        "FvTp2d": fvtp2d.FiniteVolumeTransport,
        "XPPM": xppm.XPiecewiseParabolic,
        "YPPM": yppm.YPiecewiseParabolic,
        "XTP_U": xtp_u.advect_u_along_x if False else None,
        "DelnFlux": delnflux.DelnFlux,
        "DelnFluxNoSG": delnflux.DelnFluxNoSG,
        "DivergenceDamping": divergence_damping.DivergenceDamping,
        "A2B_Ord4": a2b_ord4.AGrid2BGridFourthOrder,
        "UpdateDzD": updatedzd.UpdateHeightOnDGrid,
        "Riem_Solver3": riem_solver3.NonhydrostaticVerticalSolver,
        "PK3_Halo": pk3_halo.PK3Halo,
        "NH_P_Grad": nh_p_grad.NonHydrostaticPressureGradient,
        "Ray_Fast": ray_fast.RayleighDamping,
        "Del2Cubed": del2cubed.HyperdiffusionDamping,
        "Tracer2D1L": tracer_2d_1l.TracerAdvection,
        "Remapping": remapping.LagrangianToEulerian,
        "MapSingle": map_single.MapSingle,
        "MapNTracer": mapn_tracer.MapNTracer,
        "RemapProfile": remap_profile.RemapProfile,
        "Fillz": fillz.FillNegativeTracerValues,
        "NegAdj3": neg_adj3.AdjustNegativeTracerMixingRatio,
        "SatAdjust3d": saturation_adjustment.SatAdjust3d,
        "CubedToLatLon": c2l_ord.CubedToLatLon,
    }


ACOUSTIC_STENCILS = {
    "_p_grad_c": "PGradC",
    "_edge_pe_stencil": "PE_Halo",
    "_compute_geopotential_stencil": "ComputeGeopotential",
    "_gz_from_surface_height_and_thickness": "GzFromDelz",
    "_apply_diffusive_heating": "DiffusiveHeating",
    "_interface_pressure_from_toa_pressure_and_thickness": "PemFromDelp",
}
DYCORE_STENCILS = {
    "_fv_setup_stencil": "FVSetup",
    "_pt_to_potential_density_pt": "PtAdjust",
    "_omega_from_w": "OmegaFromW",
}

_EXTRA = {
    "C_SW": lambda s: {"delpc": s.delpc, "ptc": s.ptc},
    "AcousticDynamics": lambda s: {
        k: getattr(s, "_" + k) for k in ("gz", "zh", "pkc", "pk3", "ut", "vt", "divgd", "heat_source", "crx",
                                          "cry", "xfx", "yfx", "ws3", "wsd", "pem", "zs")
    } | {"cappa": s.cappa},
}

_patched = []


def install_capture(capture):
    for stage, cls in _stage_classes().items():
        if cls is None:
            continue
        orig = cls.__call__
        cls.__call__ = wrap_callable(orig, stage, capture, extra=_EXTRA.get(stage))
        _patched.append((cls, orig))


def uninstall_capture():
    while _patched:
        cls, orig = _patched.pop()
        cls.__call__ = orig


def build_rank(world, rank, nx, layout, config, backend="numpy", capture=None):
    _tls.rank = rank
    comm = ThreadComm(world, rank)
    partitioner = pace.util.CubedSpherePartitioner(pace.util.TilePartitioner(tuple(layout)))
    communicator = pace.util.CubedSphereCommunicator(comm, partitioner)
    with _build_lock:
        dace_config = DaceConfig(communicator=communicator, backend=backend)
        stencil_config = pace.dsl.stencil.StencilConfig(
            compilation_config=pace.dsl.stencil.CompilationConfig(backend=backend, rebuild=False, validate_args=False),
            dace_config=dace_config,
        )
    sizer = pace.util.SubtileGridSizer.from_tile_params(
        nx_tile=nx, ny_tile=nx, nz=config.npz, n_halo=3, extra_dim_lengths={}, layout=tuple(layout),
        tile_partitioner=partitioner.tile, tile_rank=communicator.tile.rank,
    )
    grid_indexing = pace.dsl.stencil.GridIndexing.from_sizer_and_communicator(sizer=sizer, cube=communicator)
    quantity_factory = pace.util.QuantityFactory.from_backend(sizer=sizer, backend=backend)
    metric_terms = MetricTerms(quantity_factory=quantity_factory, communicator=communicator)
    grid_data = GridData.new_from_metric_terms(metric_terms)
    damping = DampingCoefficients.new_from_metric_terms(metric_terms)
    state = baroclinic_init.init_baroclinic_state(
        grid_data, quantity_factory=quantity_factory, adiabatic=config.adiabatic, hydrostatic=config.hydrostatic,
        moist_phys=config.moist_phys, comm=communicator,
    )
    out = dict(rank=rank, communicator=communicator, grid_data=grid_data, damping=damping, state=state,
               metric_terms=metric_terms, quantity_factory=quantity_factory, grid_indexing=grid_indexing,
               sizer=sizer, config=config)
    return out, stencil_config


def build_dycore(ctx, stencil_config, capture=None):
    with _build_lock:
        stencil_factory = pace.dsl.stencil.StencilFactory(config=stencil_config, grid_indexing=ctx["grid_indexing"])
        dycore = fv3core.DynamicalCore(
            comm=ctx["communicator"], grid_data=ctx["grid_data"], stencil_factory=stencil_factory,
            quantity_factory=ctx["quantity_factory"], damping_coefficients=ctx["damping"], config=ctx["config"],
            timestep=timedelta(seconds=ctx["config"].dt_atmos), phis=ctx["state"].phis, state=ctx["state"],
        )
        if capture is not None:
            ad = dycore.acoustic_dynamics
            for attr, stage in ACOUSTIC_STENCILS.items():
                if hasattr(ad, attr):
                    setattr(ad, attr, wrap_callable(getattr(ad, attr), stage, capture))
            for attr, stage in DYCORE_STENCILS.items():
                if hasattr(dycore, attr):
                    setattr(dycore, attr, wrap_callable(getattr(dycore, attr), stage, capture))
    ctx["dycore"] = dycore
    ctx["stencil_factory"] = stencil_factory
    return dycore


def grid_arrays(ctx):
    """All metric terms / damping coefficients of one rank as plain numpy arrays (keyed by GridData attribute)."""
    gd = ctx["grid_data"]
    out = {}
    names = [n for n in dir(type(gd)) if isinstance(getattr(type(gd), n, None), property)]
    for n in names:
        try:
            v = getattr(gd, n)
        except Exception:
            continue
        if isinstance(v, pace.util.Quantity):
            out[n] = np.array(v.data, copy=True)
        elif isinstance(v, np.ndarray):
            out[n] = np.array(v, copy=True)
        elif isinstance(v, (int, float, np.floating, np.integer)):
            out[n] = np.asarray(v)
    d = ctx["damping"]
    for n in ("divg_u", "divg_v", "del6_u", "del6_v", "da_min", "da_min_c"):
        v = getattr(d, n)
        out["damp_" + n] = np.array(v.data if isinstance(v, pace.util.Quantity) else v, copy=True)
    return out


def state_arrays(state):
    out = {}
    for name in state.__dataclass_fields__:
        x = getattr(state, name)
        if isinstance(x, pace.util.Quantity):
            out[name] = np.array(x.data, copy=True)
    return out


TRACERS = ("qvapor", "qliquid", "qrain", "qice", "qsnow", "qgraupel", "qo3mr", "qsgs_tke")


def fill_tracers(ctx, scale=0.1):
    """BASELINE configs[3]: all 8 advected tracers non-trivial, q_m = qvapor * (m + 1) / 10 — the
    same rule as pace_b200.fv3core.initialization.baroclinic.fill_tracers, applied to the reference's own state."""
    st = ctx["state"]
    qv = np.array(st.qvapor.data, copy=True)
    for m, name in enumerate(TRACERS[1:], start=1):
        getattr(st, name).data[:] = qv * (m + 1) * scale


def run(nx, layout=(1, 1), nsteps=1, capture_ranks=(0,), stages=None, config_overrides=None, build_only=False,
        on_built=None, verbose=True, capture_step=0):
    """Run `nsteps` of the reference dycore on 6*layout^2 thread-ranks.

    Returns (contexts, capture).  contexts[r] holds grid_data/state/dycore of rank r plus
    'state0' (initial state arrays) and 'timing' (seconds per step).
    """
    config = make_config(nx, layout, **(config_overrides or {}))
    total = 6 * layout[0] * layout[1]
    world = World(total)
    capture = Capture(capture_ranks, stages)
    install_capture(capture)
    ctxs = [None] * total
    errors = []

    def work(rank):
        try:
            ctx, stencil_config = build_rank(world, rank, nx, layout, config, capture=capture)
            ctxs[rank] = ctx
            ctx["state0"] = state_arrays(ctx["state"])
            if build_only:
                return
            build_dycore(ctx, stencil_config, capture)
            if on_built is not None:
                on_built(ctx)
            ctx["timing"] = []
            for step in range(nsteps):
                world.barrier.wait()
                if rank in capture.ranks or rank == 0:
                    capture.enabled = step == capture_step
                if step == capture_step:
                    ctx["state_before_capture"] = state_arrays(ctx["state"])
                t0 = time.time()
                ctx["dycore"].step_dynamics(ctx["state"], pace.util.NullTimer())
                world.barrier.wait()
                ctx["timing"].append(time.time() - t0)
                if verbose and rank == 0:
                    print(f"[reference] step done in {ctx['timing'][-1]:.1f}s", flush=True)
        except BaseException as e:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            errors.append((rank, e))
            world.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(total)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    uninstall_capture()
    if errors:
        raise RuntimeError(f"reference run failed on ranks {[r for r, _ in errors]}: {errors[0][1]!r}")
    return ctxs, capture
