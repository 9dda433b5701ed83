"""Minimal pure-Python stand-in for the subset of toolz GT4Py uses (oracle shim only)."""
import functools
import inspect

from . import functoolz, itertoolz  # noqa: F401
from .functoolz import complement, compose, curry, identity  # noqa: F401
