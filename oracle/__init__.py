"""CPU oracle for the lbm-wgpu lattice update — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (lbm_b200/) never does.  PARITY UNPINNED: see lbm_oracle.c.
"""
