"""Minimal pure-Python stand-in for the subset of boltons GT4Py imports (oracle shim only)."""
