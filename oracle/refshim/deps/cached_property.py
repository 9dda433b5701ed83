"""Stand-in for the `cached_property` package (test infrastructure, oracle shim)."""
from functools import cached_property  # noqa: F401
