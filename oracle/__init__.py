"""CPU oracle for the mxm/mxv/vxm hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under pygraphblas_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
