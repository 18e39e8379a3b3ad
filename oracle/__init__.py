"""Test infrastructure only.

`oracle/` holds CPU restatements of the reference's VxmDense hot path (numpy and
torch-CPU).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The product package
(`voxelmorph_b200`) never imports anything from here.
"""
