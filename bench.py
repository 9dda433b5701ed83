#!/usr/bin/env python
"""Benchmark of the FV3 dycore hot path: seconds per C128 L79 baroclinic timestep (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one `DynamicalCore.step_dynamics` (k_split=2 remap cycles x n_split=6 acoustic substeps, tracer advection
of 8 tracers, vertical remap) on the analytic Jablonowski-Williamson state, C128 L79, layout (2,2) = 24 subdomains spread
over the N GPUs.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from datetime import timedelta

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# compulsory field passes (reads + writes of distinct 3-D fields) of each fused stage
STAGE_PASSES = {
    "fv3_c_sw": 15, "fv3_update_dz_c": 4, "fv3_riem_solver_c": 8, "fv3_p_grad_c": 7, "fv3_d_sw": 34,
    "fv3_update_dz_d": 6, "fv3_riem_solver3": 13, "fv3_nh_p_grad": 8, "fv3_ray_fast": 6,
    "fv3_tracer_subcycle": 25,   # one sub-cycle: 8 shared reads + dp2 write + 8 x (tracer read + write)
}
SUBSTEP_PASSES, TRACER_PASSES, REMAP_PASSES, EXTRA_BYTES_PER_CELL = 101, 85, 36, 100


def algorithmic_bytes_per_cell(k_split, n_split):
    return k_split * (n_split * SUBSTEP_PASSES * 8 + TRACER_PASSES * 8 + REMAP_PASSES * 8) + EXTRA_BYTES_PER_CELL


def build_dycore(nx, layout, nz, k_split, n_split, device, process_comm=None, all_tracers=True):
    from pace_b200.fv3core._config import baroclinic_config
    from pace_b200.fv3core.initialization import baroclinic
    from pace_b200.fv3core.runtime import Runtime
    from pace_b200.fv3core.stencil_factory import GridIndexing, StencilFactory
    from pace_b200.fv3core.stencils.fv_dynamics import DynamicalCore
    from pace_b200.util.communicator import CubedSphereCommunicator, ProcessComm
    from pace_b200.util.grid.helper import DampingCoefficients, GridData
    from pace_b200.util.sizer import QuantityFactory, SubtileGridSizer

    comm = CubedSphereCommunicator.from_layout(process_comm or ProcessComm(), (layout, layout), nx, nz, device=device)
    sizer = SubtileGridSizer.from_tile_params(nx, nx, nz, 3, {}, (layout, layout))
    qf = QuantityFactory(sizer, comm.geometry, device)
    gd = GridData.new_from_generation(qf, comm)
    damp = DampingCoefficients.new_from_generation(qf, gd)
    cfg = baroclinic_config(nx, (layout, layout), n_split=n_split, k_split=k_split)
    rt = Runtime(comm, qf, gd, damp, cfg)
    sf = StencilFactory(None, GridIndexing.from_sizer_and_communicator(qf.sizer, comm), rt)
    state = baroclinic.init_baroclinic_state(gd, qf, adiabatic=False, hydrostatic=False, moist_phys=True, comm=comm,
                                             fill_all_tracers=all_tracers)
    dycore = DynamicalCore(comm, gd, sf, qf, damp, cfg, state.phis, state, timedelta(seconds=cfg.dt_atmos))
    return dycore, state, comm, rt, gd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_run(nx, layout, k_split, n_split, steps, warmup, budget_s=None):
    """The CPU port (host simulation of the same stage functors, OpenMP over the host cores; oracle/hostsim.py) on the
    C`nx` workload: (mean seconds per step, measured steps, threads actually used).  With `budget_s` the number of
    measured steps is cut (never below 1) so that the whole run fits the budget."""
    import torch

    from oracle import hostsim

    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)   # unconditional: torchrun exports OMP_NUM_THREADS=1
    lib = hostsim.install(openmp=True, threads=cores)
    threads = hostsim.threads_in_use(lib)
    dycore, state, comm, rt, gd = build_dycore(nx, layout, 79, k_split, n_split, "cpu")
    t0 = time.perf_counter()
    dycore.step_dynamics(state)          # first step: also the calibration of the budget
    t_first = time.perf_counter() - t0
    n_warm = max(warmup - 1, 0)
    if budget_s is not None:
        room = max(int(budget_s / max(t_first, 1e-3)) - 1, 1)
        n_warm = min(n_warm, max(room - steps, 0))
        steps = max(min(steps, room - n_warm), 1)
    for _ in range(n_warm):
        dycore.step_dynamics(state)
    t0 = time.perf_counter()
    for _ in range(steps):
        dycore.step_dynamics(state)
    dt = (time.perf_counter() - t0) / steps
    assert bool(torch.isfinite(state.pt.data[:, 3:-4, 3:-4, :79]).all()), "CPU port produced non-finite values"
    return dt, steps, n_warm + 1, threads


def reference_numpy_note():
    """The reference's OWN numpy backend cannot run on the GPU box (Python + vendored GT4Py, /root/reference is not
    shipped); its measured cost in the build container is recorded by oracle/refshim/time_reference.py in
    profiles/reference_numpy_timing.json and repeated here for orientation."""
    p = os.path.join(ROOT, "profiles", "reference_numpy_timing.json")
    return json.load(open(p)) if os.path.exists(p) else None


def run_reference_arm(args):
    """`--impl reference`: the CPU implementation of the path on the host cores.  The reference itself is Python + a
    GT4Py tool-chain that is neither installable nor present on the GPU box, so this arm times the CPU port
    (oracle/hostsim.py: this repo's stage functors compiled for the host, OpenMP) — `kind: "port"`, NOT a measurement
    of the reference's code — on the SAME workload (C128 by default), a bounded number of steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, steps, warm, threads = cpu_port_run(args.nx, args.layout, args.k_split, args.n_split, args.steps, args.warmup, budget_s=240.0)
    sample = (f"C{args.nx} L79 layout ({args.layout},{args.layout}) full timestep (k_split={args.k_split}, n_split={args.n_split}), the workload "
              f"itself: mean of {steps} step(s) after {warm} warm-up step(s) on {threads} OpenMP threads (steps cut to a 240 s budget)")
    line = {
        "impl": "reference", "metric": "C128L79 baroclinic dycore s/timestep", "value": value, "unit": "s/timestep",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": value * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "s/timestep", "cores": threads, "kind": "port", "sample": sample,
                         "reference_numpy": reference_numpy_note()},
        "e2e": {"value": value, "unit": "s/timestep", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def parity_block(device):
    """Strict one-timestep parity of the benchmarked split (c12 L79, k_split=2, n_split=6, 8 non-zero tracers) against the
    reference's final state (tests/step_cases.py, fixtures tests/golden/c12k2n6_step): the achieved worst relative error
    (over the points above the absolute floor) and worst absolute error per prognostic field, measured on this GPU."""
    try:
        from tests import step_cases as S

        if not S.available("c12k2n6"):
            return {"case": "c12k2n6", "status": "fixtures not present"}
        meta, grids, s0, s1 = S.load("c12k2n6")
        dycore, state = S.build(meta, grids, s0, device)
        dycore.step_dynamics(state)
        import torch

        torch.cuda.synchronize()
        failures, achieved = S.compare(state.as_numpy(), s1, meta)
        return {"case": "c12 L79 layout (1,1), k_split=2, n_split=6, 8 non-zero tracers, reference = unmodified ai2cm/pace numpy backend",
                "status": "pass" if not failures else f"{len(failures)} failures", "tolerance": "relative 1e-10 or the reference's calibrated floors",
                "worst_rel": {k: v[0] for k, v in achieved.items()}, "worst_abs": {k: v[1] for k, v in achieved.items()}}
    except Exception as exc:  # the bench line must not die on the side check
        return {"case": "c12k2n6", "status": f"not run: {exc!r}"}


def fp64_peak_tflops(lib, device):
    """Measured fp64 FMA throughput of this GPU (fv3_fp64_peak: 8 independent FMA chains per thread on every SM): the
    instruction roof of the plane / column kernels, which ncu shows to be issue- and fp64-bound rather than HBM-bound."""
    import torch

    blocks, iters = 148 * 16, 20000
    out = torch.empty(blocks * 128, dtype=torch.float64, device=device)
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.fv3_fp64_peak(out.data_ptr(), iters, blocks, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return 2.0 * 8 * iters * 128 * blocks / (best / 1e3) / 1e12


def workload_config(args):
    return {
        "workload": f"C{args.nx} L79 baroclinic (Jablonowski-Williamson), layout ({args.layout},{args.layout}) = "
                    f"{6 * args.layout ** 2} subdomains, DynamicalCore.step_dynamics, k_split={args.k_split}, n_split={args.n_split}, "
                    f"dt_atmos=225, 8 non-zero tracers, do_sat_adj off",
        "nx_tile": args.nx, "nz": 79, "layout": [args.layout, args.layout], "k_split": args.k_split, "n_split": args.n_split,
        "subdomains_per_gpu": 6 * args.layout ** 2 // max(args.gpus, 1),
        "l2_policy": "working set (>1 GB of fp64 fields per GPU) exceeds the 126 MB L2; no explicit flush",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=128)
    ap.add_argument("--layout", type=int, default=2)
    ap.add_argument("--k-split", type=int, default=2)
    ap.add_argument("--n-split", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--stage-table", action="store_true", help="print the per-stage timing table to stderr")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of the captured CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: raising --warmup to the required minimum of 3", file=sys.stderr)
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch

    from pace_b200 import _lib
    from pace_b200.util.communicator import ProcessComm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    pc = None
    dist = None
    if world > 1:
        import torch.distributed as dist

        # rank 0 prints ONE JSON line on stdout: NCCL's version banner / debug output goes to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(device))
        pc = ProcessComm.from_torch_distributed()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib = _lib.load()
    parity = parity_block(device) if rank == 0 else None
    fp64_peak = fp64_peak_tflops(lib, device) if rank == 0 else None
    dycore, state, comm, rt, gd = build_dycore(args.nx, args.layout, 79, args.k_split, args.n_split, device, pc)
    n_local = comm.geometry.n_sub
    cells_local = n_local * comm.geometry.nx * comm.geometry.ny * 79
    cells_total = 6 * args.nx * args.nx * 79
    from pace_b200.fv3core.dycore_state import FIELDS

    def storage(q):
        """The contiguous allocation behind a Quantity (I-fastest, padded rows) — what is copied to / from the host."""
        d = q.data
        return d._base if d._base is not None else d

    state0 = {n: storage(getattr(state, n)).clone() for n in FIELDS}

    # ---- eager pass: per-stage CUDA-event timing (roofline), launch count, host enqueue time ----------------
    for _ in range(args.warmup):
        dycore.step_dynamics(state)
    barrier()
    prof = _lib.PROFILE = _lib.StageProfile()
    l0 = lib.fv3_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host = time.perf_counter()
    n_eager = args.steps if args.no_graph else min(args.steps, 2)
    for _ in range(n_eager):
        dycore.step_dynamics(state)
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / n_eager
    e1.record()
    barrier()
    launches = (lib.fv3_launch_count() - l0) * args.steps // n_eager
    _lib.PROFILE = None
    ms_eager = e0.elapsed_time(e1) / n_eager
    ms = ms_eager
    mode = "eager"
    clocks = None
    # ---- device-resident timing (value): K replays of the captured timestep graph -----------------------------
    if not args.no_graph:
        try:
            replay = dycore.capture_step(state)
            for _ in range(2):
                replay()
            barrier()
            sampler = ClockSampler(local_rank) if rank == 0 else None
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            g0.record()
            for _ in range(args.steps):
                replay()
            g1.record()
            barrier()
            ms = g0.elapsed_time(g1) / args.steps
            clocks = sampler.stop() if sampler is not None else None
            mode = "cuda_graph"
        except Exception as exc:  # capture unsupported (e.g. a library op that cannot be captured): keep eager
            print(f"bench.py: CUDA graph capture failed, timing eager launches instead: {exc!r}", file=sys.stderr)
            args.no_graph = True
    if args.no_graph:
        sampler = ClockSampler(local_rank) if rank == 0 else None
        barrier()
        e0.record()
        for _ in range(args.steps):
            dycore.step_dynamics(state)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        clocks = sampler.stop() if sampler is not None else None
    finite = bool(torch.isfinite(state.pt.data[:, 3:-4, 3:-4, :79]).all())
    # per-subdomain sums of a few prognostic fields over the compute domain, gathered in subdomain order: identical
    # for every N because kernels are deterministic per subdomain and the halo exchange is pure data movement
    sums = torch.stack([torch.stack([getattr(state, n).data[s_, 3:-4, 3:-4, :79].sum() for n in ("u", "w", "delp", "pt", "qvapor")])
                        for s_ in range(n_local)])
    if dist is not None:
        parts = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(parts, sums)
        sums = torch.cat(parts)
    import hashlib

    digest = hashlib.sha1(sums.cpu().numpy().tobytes()).hexdigest()[:16]
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    stages = prof.summary()

    # ---- end to end through the public API with HOST buffers (e2e) --------------------------------------------
    e2e = None
    if not args.no_e2e:
        names = list(state0)
        host_in = {n: state0[n].cpu().pin_memory() for n in names}
        host_out = {n: torch.empty_like(host_in[n]).pin_memory() for n in names}
        h2d = sum(v.numel() * 8 for v in host_in.values())
        d2h = h2d

        def e2e_step():
            for n in names:
                storage(getattr(state, n)).copy_(host_in[n], non_blocking=True)
            if mode == "cuda_graph":
                replay()
            else:
                dycore.step_dynamics(state)
            for n in names:
                host_out[n].copy_(storage(getattr(state, n)), non_blocking=True)

        for _ in range(2):
            e2e_step()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            e2e_step()
        a1.record()
        barrier()
        t = torch.tensor([a0.elapsed_time(a1) / n_e2e], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": float(t.item()) / 1e3, "unit": "s/timestep", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "steps": n_e2e, "note": "full DycoreState copied from pinned host memory before and back to it after every step_dynamics"}

    if rank != 0:
        return finish(dist)

    # ---- roofline of the dominant stage -------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    table = sorted(((t_ms, n, name) for name, (n, t_ms) in stages.items()), reverse=True)
    if args.stage_table:
        for t_ms, n, name in table:
            per = t_ms / n
            extra = ""
            if name in STAGE_PASSES:
                gb = STAGE_PASSES[name] * 8 * cells_local / 1e9
                extra = f"  {gb / (per / 1e3):8.1f} GB/s algorithmic"
            print(f"{name:32s} calls {n:5d}  total {t_ms:9.3f} ms  avg {per * 1e3:9.1f} us{extra}", file=sys.stderr)
    dom = next(((t_ms, n, name) for t_ms, n, name in table if name in STAGE_PASSES), None)
    roofline = None
    if dom is not None:
        t_ms, n, name = dom
        bytes_per_launch = STAGE_PASSES[name] * 8 * cells_local
        achieved = bytes_per_launch / (t_ms / n / 1e3) / 1e9
        traffic, fp64 = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get(name)
            # fp64 roof of the same stage: DADD + DMUL + 2 x DFMA thread-instructions per call from the same ncu pass
            # (C128, all 24 subdomains), scaled to the cells this rank owns
            ops = tj.get("_fp64_flops", {}).get(name)
            if ops and fp64_peak:
                flops = ops * cells_local / (6 * args.nx * args.nx * 79)
                fp64 = {"flops_per_launch": flops, "achieved_tflops": flops / (t_ms / n / 1e3) / 1e12, "peak_tflops": fp64_peak,
                        "frac": flops / (t_ms / n / 1e3) / 1e12 / fp64_peak,
                        "peak_source": "measured in this run (fv3_fp64_peak, 8 independent FMA chains per thread)"}
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "fp64": fp64, "binding_roof": ("fp64" if fp64 and fp64["frac"] > achieved / peak else "hbm"),
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                    "avg_launch_ms": t_ms / n, "share_of_step": t_ms / (ms_eager * n_eager),
                    "step_achieved_GBps": algorithmic_bytes_per_cell(args.k_split, args.n_split) * cells_total / (ms / 1e3) / 1e9,
                    "step_frac": algorithmic_bytes_per_cell(args.k_split, args.n_split) * cells_total / (ms / 1e3) / 1e9 / (peak * world)}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        _prod = _lib._lib
        try:
            dt, n_meas, n_warm, threads = cpu_port_run(args.nx, args.layout, args.k_split, args.n_split, 1, 1, budget_s=30.0)
        finally:
            _lib.install(_prod)   # back to the CUDA library
        cpu = {"value": dt, "unit": "s/timestep", "cores": threads, "kind": "port",
               "sample": f"C{args.nx} L79 layout ({args.layout},{args.layout}) full timestep (same k_split/n_split), the workload itself: "
                         f"{n_meas} step after {n_warm} warm-up on {threads} OpenMP threads",
               "reference_numpy": reference_numpy_note()}

    line = {
        "metric": "C128L79 baroclinic dycore s/timestep", "value": ms / 1e3, "unit": "s/timestep", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args), "clocks": clocks,
        "halo_transport": ("device-local gathers only" if world == 1 else
                           ("fv3_halo_exchange_nccl (C ABI, own ncclComm_t)" if pc.nccl_comm is not None else "torch.distributed.batch_isend_irecv")),
        # per-stage device time of the eager pass (CUDA events around every C-ABI call), per timestep, largest first
        "stages_ms_per_step": {name: round(t_ms / n_eager, 4) for t_ms, n, name in table[:14]},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "fp64_peak_tflops": fp64_peak, "finite": finite,
        "state_digest": digest,
        "launch_mode": mode, "ms_per_step_eager": ms_eager, "host_enqueue_ms_per_step": host_enqueue_ms,
    }
    print(json.dumps(line), flush=True)
    finish(dist)


def finish(dist):
    """End the process once its work is done.  Multi-rank runs leave through os._exit: tearing down an NCCL process group
    whose send/recv kernels were captured into a CUDA graph can block in the communicator's finalizer, and nothing
    remains to be flushed but the standard streams."""
    sys.stdout.flush()
    sys.stderr.flush()
    if dist is not None:
        import torch

        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
