#!/bin/bash
# GPU session AE (round 1): fuzz soak — 300 extra random API walks (half of them on a 15 % porous mask) against
# the oracle, every kernel / flavour / knob toggled at random, bit-exact comparison after every few operations.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time BLBM_FUZZ_SEEDS=1000-1299 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -n 4 -k "fuzz_against" ) 2>&1 | tail -8 | tee gpurun_out/ae_fuzz_soak.log
