"""Oracle shim: imports the UNMODIFIED reference (ai2cm/pace, /root/reference) under Python 3.12 / numpy 2.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/refshim/gen_golden.py` (golden-vector generation) and by the
container-only tests that validate the numpy restatement in `oracle/` against the real reference.
Nothing under `pace_b200/` may import this package, and nothing here runs on the GPU box
(`/root/reference` does not exist there).
"""
