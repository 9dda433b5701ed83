"""CPU host-simulation of the kernel sources — TEST INFRASTRUCTURE, never part of the product path.

The stage functors of pace_b200/csrc/*.cu are written against launch3d/launch2d/launch1d (csrc/common.h); compiled
with g++ -DFV3_HOSTSIM those launchers become plain loops, which lets CPU-only CI exercise the host-side orchestration
(stage order, halo tables, multi-process exchange over gloo) and gives bench.py a multi-threaded CPU port to time as
the `cpu_baseline` / `--impl reference` arm (OpenMP over the launch loops).  Only tests/, __graft_entry__.smoke() and
bench.py's CPU legs may import this module; pace_b200 itself can only load the CUDA library (pace_b200/_lib.py).
What pins numerical parity is NOT this build but the reference golden vectors in tests/golden.
"""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "oracle", "_hostsim")
GXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-DFV3_HOSTSIM", "-x", "c++", "-w"]


def lib_path(openmp=False):
    return os.path.join(OUT, "libfv3b200_hostsim_omp.so" if openmp else "libfv3b200_hostsim.so")


def build(openmp=False, verbose=False):
    from pace_b200 import build as b

    gxx = os.environ.get("FV3_HOSTSIM_CXX", "g++")
    flags = GXX_FLAGS + (["-fopenmp", "-DFV3_HOSTSIM_OMP", "-O3"] if openmp else [])
    objdir = os.path.join(OUT, "obj_omp" if openmp else "obj")
    objs, changed, _ = b.compile_all(lambda s, o: [gxx] + flags + ["-c", s, "-o", o], objdir, verbose)
    target = lib_path(openmp)
    if changed or not os.path.exists(target):
        extra = ["-fopenmp"] if openmp else []
        r = subprocess.run([gxx, "-shared", "-o", target] + objs + extra, capture_output=True, text=True)
        if r.returncode != 0 and openmp:  # toolchains without libgomp.spec: link the runtime directly
            r = subprocess.run([gxx, "-shared", "-o", target] + objs + ["-l:libgomp.so.1"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return target


def load(openmp=False, build_if_missing=True):
    """Bound ctypes library of the host simulation (built on demand when a compiler is present)."""
    from pace_b200 import _lib

    path = lib_path(openmp)
    if build_if_missing:
        try:
            path = build(openmp)
        except Exception:
            if not os.path.exists(path):
                raise
    lib = _lib.bind(ctypes.CDLL(path))
    assert lib.fv3_is_hostsim() == 1
    return lib


def install(openmp=False, threads=None):
    """Make pace_b200 run on the host simulation in THIS process (tests / CPU baseline only).  `threads`: OpenMP
    thread count set EXPLICITLY through the runtime (an inherited OMP_NUM_THREADS=1, as torchrun exports, would
    otherwise pin the port to one core)."""
    from pace_b200 import _lib

    lib = load(openmp)
    if openmp and threads:
        lib.omp_set_num_threads(int(threads))
    _lib.install(lib)
    return lib


def threads_in_use(lib):
    """OpenMP threads the host-simulation library will actually use (1 for the serial build)."""
    try:
        return int(lib.omp_get_max_threads())
    except AttributeError:
        return 1
