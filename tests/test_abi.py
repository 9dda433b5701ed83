"""The drop-in boundary: pace_b200/libfv3b200.so (CUDA, sm_100a) loads and exports every symbol include/fv3_b200.h
declares, and the product path fails loudly — it never falls back to a CPU implementation — when the library is
missing.  No compute call is made, so this runs without a GPU."""
import ctypes
import os

import pytest

from pace_b200 import _lib, build

EXTRA = ["fv3_create", "fv3_destroy", "fv3_last_error", "fv3_is_hostsim", "fv3_abi_version", "fv3_scratch_fields",
         "fv3_launch_count"]


@pytest.fixture(scope="module")
def cuda_lib():
    lib_path, _ = build.build_cuda()   # no-op when the in-tree library is up to date
    assert os.path.exists(lib_path)
    return ctypes.CDLL(lib_path)


def test_header_declares_the_hot_path(cuda_lib):
    decl = _lib.parse_header()
    for stage in ("fv3_c_sw", "fv3_d_sw", "fv3_riem_solver_c", "fv3_riem_solver3", "fv3_nh_p_grad", "fv3_update_dz_c",
                  "fv3_update_dz_d", "fv3_tracer_subcycle", "fv3_map_single", "fv3_map_multi", "fv3_fillz",
                  "fv3_halo_gather", "fv3_halo_pack", "fv3_halo_unpack", "fv3_neg_adj3", "fv3_fvtp2d"):
        assert stage in decl, stage


def test_library_exports_every_declared_symbol(cuda_lib):
    missing = [n for n in list(_lib.parse_header()) + EXTRA if not hasattr(cuda_lib, n)]
    assert not missing, missing
    cuda_lib.fv3_is_hostsim.restype = ctypes.c_int
    cuda_lib.fv3_abi_version.restype = ctypes.c_int
    assert cuda_lib.fv3_is_hostsim() == 0          # the product library is the CUDA build
    assert cuda_lib.fv3_abi_version() == 1
    _lib.bind(cuda_lib)                            # every prototype of the header binds (argtypes parse)


def test_missing_library_is_a_loud_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "HERE", str(tmp_path))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_package_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    offenders = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(d, f)).read()
                if "import oracle" in text or "from oracle" in text:
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders
