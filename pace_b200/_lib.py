"""ctypes binding of libfv3b200.so (include/fv3_b200.h).

The product path has NO CPU fallback: `load()` only ever loads the CUDA library pace_b200/libfv3b200.so and raises
if it has not been built.  (CPU-only CI of the host-side orchestration injects a host-simulation build of the same
kernel sources through `install()`; that build and its loader live under oracle/hostsim.py — test infrastructure —
and nothing in this package can reach them.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_SUB = 64

c_dp = C.POINTER(C.c_double)


class Geom(C.Structure):
    _fields_ = [
        ("n_sub", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("halo", C.c_int32),
        ("ni", C.c_int32), ("nj", C.c_int32), ("nk", C.c_int32), ("sj", C.c_int32), ("pad_", C.c_int32),
        ("sk", C.c_int64), ("ss", C.c_int64), ("ss2", C.c_int64), ("edge", C.c_uint8 * MAX_SUB),
    ]


CONFIG_INT = ["hord_dp", "hord_tm", "hord_mt", "hord_vt", "hord_tr", "kord_tm", "kord_tr", "kord_wz", "kord_mt",
              "nord", "n_sponge", "nwat", "fill", "do_vort_damp", "convert_ke", "hydrostatic", "rf_fast", "ks"]
CONFIG_DBL = ["d2_bg", "d2_bg_k1", "d2_bg_k2", "d4_bg", "ke_bg", "dddmp", "vtdm4", "d_con", "delt_max", "p_fac",
              "a_imp", "tau", "rf_cutoff", "ptop", "da_min", "da_min_c"]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in CONFIG_INT] + [(n, C.c_double) for n in CONFIG_DBL]


GRID_FIELDS = [
    "dx", "dy", "dxa", "dya", "dxc", "dyc", "rdx", "rdy", "rdxa", "rdya", "rdxc", "rdyc",
    "area", "area_64", "rarea", "rarea_c",
    "cosa", "cosa_u", "cosa_v", "cosa_s", "sina_u", "sina_v", "rsina", "rsin_u", "rsin_v", "rsin2",
    "sin_sg1", "sin_sg2", "sin_sg3", "sin_sg4", "cos_sg1", "cos_sg2", "cos_sg3", "cos_sg4",
    "fC", "f0",
    "edge_w", "edge_e", "edge_s", "edge_n",
    "divg_u", "divg_v", "del6_u", "del6_v",
    "a11", "a12", "a21", "a22",
    "ak", "bk", "dp_ref", "pfull",
    "a2b_w",
]


class Grid(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in GRID_FIELDS]


DSW_COLS_PTR = ["nord", "nord_v", "nord_w", "nord_t", "damp_vt", "damp_w", "damp_t", "d_con", "ke_bg", "d2_divg",
                "dn_damp_vt", "dn_damp_t", "dn_damp_vt_c", "dn_damp_w_c"]
DSW_COLS_INT = ["nmax_v", "nmax_w", "nmax_t", "nonzero_nord_k", "nonzero_nord"]


class DswCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in DSW_COLS_PTR] + [(n, C.c_int32) for n in DSW_COLS_INT]


_lib = None
HEADER = os.path.join(os.path.dirname(HERE), "include", "fv3_b200.h")


def parse_header(path=HEADER):
    """{function name: ctypes argtypes} for every `int fv3_*(...)` prototype declared in include/fv3_b200.h."""
    import re

    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out = {}
    for m in re.finditer(r"\bint\s+(fv3_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append(C.c_void_p)
                elif re.match(r"(const\s+)?double\b", a):
                    types.append(C.c_double)
                elif re.match(r"(const\s+)?int64_t\b", a):
                    types.append(C.c_int64)
                elif re.match(r"(const\s+)?(int|int32_t)\b", a):
                    types.append(C.c_int)
                else:
                    raise ValueError(f"unhandled C type in {name}: {a}")
        out[name] = types
    return out



def bind(lib):
    """Declare restype/argtypes of every entry point of include/fv3_b200.h on a loaded ctypes library."""
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.fv3_create.restype = vp
    lib.fv3_create.argtypes = [C.POINTER(Geom), C.POINTER(Config), C.POINTER(Grid), vp, i64]
    lib.fv3_destroy.argtypes = [vp]
    lib.fv3_last_error.restype = C.c_char_p
    lib.fv3_is_hostsim.restype = i32
    lib.fv3_abi_version.restype = i32
    lib.fv3_scratch_fields.restype = i32
    lib.fv3_launch_count.restype = i64
    for name, argtypes in parse_header().items():
        if name in ("fv3_create", "fv3_destroy"):
            continue
        fn = getattr(lib, name)  # AttributeError here = header declares a symbol the library lacks
        fn.restype = i32
        fn.argtypes = argtypes
    return lib


def load():
    """Load (once) and return the CUDA library.  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(HERE, "libfv3b200.so")
    if not os.path.exists(path):
        raise RuntimeError(
            f"pace_b200: native library {path} not found; build it with `python -m pace_b200.build` "
            "(there is no CPU fallback for the hot path)"
        )
    lib = bind(C.CDLL(path))
    if lib.fv3_is_hostsim():
        raise RuntimeError("pace_b200: libfv3b200.so is not a CUDA build")
    _lib = lib
    return lib


def install(lib):
    """Use an already bound library object instead of loading libfv3b200.so (tests only: oracle/hostsim.py)."""
    global _lib
    _lib = lib


class StageProfile:
    """CUDA-event timing of every C-ABI stage call on the launching stream (bench.py's roofline pass).

    Switched on with `_lib.PROFILE = StageProfile()`; off (None) by default so the hot path records nothing.
    """

    def __init__(self):
        self.records = []   # (stage name, start event, end event)

    def begin(self):
        import torch

        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, name, e0):
        self.records.append((name, e0, self.begin()))

    def summary(self):
        """{stage: (calls, total ms)} — call after torch.cuda.synchronize()."""
        out = {}
        for name, e0, e1 in self.records:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + e0.elapsed_time(e1))
        return out


PROFILE = None


def check(lib, rc, what=""):
    if rc != 0:
        raise RuntimeError(f"pace_b200 native call failed ({what}, rc={rc}): {lib.fv3_last_error().decode()}")
