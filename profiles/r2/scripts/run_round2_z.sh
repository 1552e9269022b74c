#!/bin/bash
# GPU session Z (1 GPU): the group-handle walk that failed once in session Y's soak (400 seeds, 4 xdist workers sharing the GPU):
# same command with the full failure report, then the failed seeds again serially.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export BLBM_FUZZ_SLAB_SEEDS=20001-20400
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -n 4 --tb=long -rf -k "api_fuzz_slab_group" > gpurun_out/r2z_group_xdist.log 2>&1
tail -5 gpurun_out/r2z_group_xdist.log
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --lf --tb=long -rf -k "api_fuzz_slab_group" > gpurun_out/r2z_group_lastfailed_serial.log 2>&1
tail -5 gpurun_out/r2z_group_lastfailed_serial.log
exit 0
