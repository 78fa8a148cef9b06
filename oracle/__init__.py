"""CPU oracle for the beta-cores hot path -- TEST INFRASTRUCTURE ONLY.

A numpy (IEEE fp64) restatement of the reference algorithm
(dionman/beta-cores, `bayesiancoresets/` + `examples/common/`).  Every function
cites the reference file:line whose arithmetic it follows, operation by
operation, so that on the same numpy build it reproduces the reference bit for
bit.  It is pinned against the reference itself: `tests/golden/make_golden.py`
imports the unmodified reference from /root/reference (through two sys.modules
shims) and writes the fixtures in `tests/golden/*.npz`; `tests/test_oracle_golden.py`
replays the oracle against those fixtures.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`beta-cores_b200/`) never does: it has no CPU path and fails loudly if the
CUDA library is missing.
"""
