#!/bin/bash
# GPU session Z4 (1 GPU): the public stream half-step publishes an epoch too (same class of hole as the summary's): the linked-slab /
# group / multi-process tests incl. the extended skew test, and 100 group walks under 4 time-slicing workers.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests -m gpu -q -x -k "running_ahead or group or slab or multiproc or handshake or linked or peer" ) 2>&1 | tail -8 | tee gpurun_out/r2z4_gpu_linked.log
( time BLBM_FUZZ_SLAB_SEEDS=20201-20300 timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -n 4 -rf -k "api_fuzz_slab_group" ) 2>&1 | tail -6 | tee gpurun_out/r2z4_fuzz_group_100_xdist.log
exit 0
