"""CPU oracle for the lbm-wgpu lattice update — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (lbm_b200/) never does.  Parity status (details in lbm_oracle.c): pinned bit for bit to
the reference's WGSL text as executed by wgsl_interp.py / wgsl_simt.py and to host-side functions of its shipped wasm
binary as executed by wasm_mini.py (the -DLBM_CONTRACT twin: to the same WGSL text executed under the fma contraction
that wgsl_contract.py derives from it); unpinned only against a run on a real WebGPU backend (none available here).
"""
