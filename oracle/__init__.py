"""CPU oracle of the FV3 hot path — TEST INFRASTRUCTURE, not product code.

A plain numpy restatement of the reference's algorithm (ai2cm/pace), one function per reference stage, each citing
the reference file:line it follows, operating on ONE subdomain's arrays in the reference's own storage order
[i, j, k] (halo 3, shape (nx+7, ny+7, nz+1)).  GT4Py semantics are kept: each statement applies to its whole
domain before the next one starts.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.
Pinning: every function here is checked against arrays dumped from the UNMODIFIED reference (numpy backend, run
through `oracle/refshim`) at its own call site — tests/golden/ (committed subset) and tests/test_oracle_*.py.
"""
