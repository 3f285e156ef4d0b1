"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the UniRec sequential-recommender hot path.

Nothing in the product package (`unirec_b200/`) may import this package.  Only
`tests/`, `__graft_entry__.smoke()` and the CPU-baseline / `--impl reference`
legs of `bench.py` use it, and only as the checker or the timed CPU baseline.

Parity status: PINNED.  `oracle/make_golden.py` imports the reference's own
model classes from /root/reference (in the build container), runs them on
seeded inputs and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py`
checks every function here against those fixtures.
"""
