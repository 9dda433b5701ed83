"""Build an experimental variant of libfv3b200.so with extra nvcc flags:  python tools/build_variant.py NAME -DFOO=1 ...
The library goes to variants/NAME.so (git-ignored); tools/variants.sh benches a list of them on the GPU box."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pace_b200 import build as b  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
objdir = os.path.join(ROOT, "variants", "obj_" + name)
nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
objs, _, outs = b.compile_all(lambda s, o: [nvcc] + b.NVCC_FLAGS + flags + ["-c", s, "-o", o], objdir)
out = os.path.join(ROOT, "variants", name + ".so")
subprocess.run([nvcc, "-shared", "-o", out] + objs + ["-lcudart"], check=True)
print(out)
