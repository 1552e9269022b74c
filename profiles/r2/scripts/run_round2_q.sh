#!/bin/bash
# GPU session Q (2 GPUs): long soak of the random API walks on the final library — 3000 single-handle walks and 1000
# walks over group handles of 2-4 slabs (first two slabs on distinct devices).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time BLBM_FUZZ_SLAB_SEEDS=7001-8000 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "api_fuzz_slab_group" ) 2>&1 | tail -6 | tee gpurun_out/r2q_fuzz_group_1000.log
( time BLBM_FUZZ_SEEDS=10001-13000 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "api_fuzz_against_oracle" ) 2>&1 | tail -6 | tee gpurun_out/r2q_fuzz_single_3000.log
exit 0
