"""Build recipe for libfv3b200.so (CUDA, sm_100a).

    python -m pace_b200.build            # nvcc cross-compiles without a GPU

The CUDA library is built IN-TREE (pace_b200/libfv3b200.so) so it travels with the repo snapshot.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
CUDA_LIB = os.path.join(HERE, "libfv3b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-fmad=" + os.environ.get("FV3_FMAD", "false"),  # fp64 parity with the numpy backend: no FMA contraction (DESIGN.md "parity budget")
    "-Xcompiler", "-fPIC",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    return glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h"))


def compile_all(cmd_for, objdir, verbose=False):
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if _stale(obj, [src] + _headers()):
            jobs.append(cmd_for(src, obj))

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stdout + r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        outs = list(ex.map(run, jobs))
    return objs, bool(jobs), outs


def build_cuda(verbose=False, extra_flags=()):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build", "cuda")
    objs, changed, outs = compile_all(
        lambda s, o: [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", s, "-o", o], objdir, verbose)
    if changed or not os.path.exists(CUDA_LIB):
        cmd = [nvcc, "-shared", "-o", CUDA_LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return CUDA_LIB, outs


if __name__ == "__main__":
    v = "-v" in sys.argv
    lib, outs = build_cuda(v, ["-Xptxas", "-v"] if "--ptxas" in sys.argv else [])
    for o in outs:
        if o.strip():
            print(o)
    print(lib)
