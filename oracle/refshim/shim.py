"""Import shim for the reference (recipe: SURVEY.md Appendix A).  Import this module FIRST.

Stubs the third-party packages the reference imports but the numpy-backend path never executes
(dace, cftime, xarray, zarr, f90nml, dacite, netCDF4, serialbox) and puts pure-Python stand-ins for
boltons / toolz / frozendict / ... (see ./deps) on sys.path.  The reference sources are used where
they lie under /root/reference; nothing is copied.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PACE_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fv3core"))


class _Any:
    """Permissive attribute sink used for every attribute of a stubbed module."""

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Any()

    def __call__(self, *a, **k):
        if len(a) == 1 and not isinstance(a[0], _Any) and not k:
            return a[0]
        return _Any()

    def __getitem__(self, k):
        return _Any()

    def __mro_entries__(self, bases):
        return (object,)

    def __or__(self, o):
        return self

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return False

    def __add__(self, o):
        return self

    __radd__ = __sub__ = __rsub__ = __add__


STUB_ROOTS = ("dace", "cftime", "xarray", "zarr", "f90nml", "dacite", "netCDF4", "serialbox")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []

        def _getattr(k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Any()

        m.__getattr__ = _getattr
        return m

    def exec_module(self, m):
        pass


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy.lib  # noqa: F401

    it = types.ModuleType("numpy.lib.index_tricks")
    it.IndexExpression = type(np.index_exp)
    sys.modules["numpy.lib.index_tricks"] = it
    for a, b in [("float_", "float64"), ("int_", "int64"), ("bool", "bool_")]:
        if not hasattr(np, a):
            setattr(np, a, getattr(np, b))
    sys.path[:0] = [
        os.path.join(HERE, "deps"),
        os.path.join(REFERENCE_ROOT, "external/gt4py/src"),
        os.path.join(REFERENCE_ROOT, "util"),
        os.path.join(REFERENCE_ROOT, "dsl"),
        os.path.join(REFERENCE_ROOT, "stencils"),
        os.path.join(REFERENCE_ROOT, "fv3core"),
        os.path.join(REFERENCE_ROOT, "driver"),
        os.path.join(REFERENCE_ROOT, "physics"),
    ]
    os.environ.setdefault("GT_CACHE_ROOT", "/tmp/pace_ref_gtcache")
    os.makedirs(os.environ["GT_CACHE_ROOT"], exist_ok=True)
    # real gt4py must be imported BEFORE dace is stubbed so that its dace backends are skipped
    import gt4py  # noqa: F401
    import gt4py.cartesian.gtscript  # noqa: F401

    sys.meta_path.insert(0, _Finder())
    _installed = True


install()
