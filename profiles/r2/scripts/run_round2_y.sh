#!/bin/bash
# GPU session Y (1 GPU): soak of the random API walks on the library as committed (one-launch paints, graph tails, dependent
# launches among the knobs the walk toggles): 3000 single-handle walks, 400 walks over group handles of 2-4 slabs.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time BLBM_FUZZ_SEEDS=20001-23000 timeout 560 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -n 4 -k "api_fuzz_against_oracle" ) 2>&1 | tail -6 | tee gpurun_out/r2y_fuzz_single_3000.log
( time BLBM_FUZZ_SLAB_SEEDS=20001-20400 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -n 4 -k "api_fuzz_slab_group" ) 2>&1 | tail -6 | tee gpurun_out/r2y_fuzz_group_400.log
exit 0
