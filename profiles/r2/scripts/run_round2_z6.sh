#!/bin/bash
# GPU session Z6 (1 GPU, the round's last GPU seconds): blbm_exchange_halos publishes an epoch before its pushes (a race found by the
# CPU protocol model): the linked-slab / group / multi-process tests once more.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 120 python -m pytest tests -m gpu -q -x -k "running_ahead or group or slab or multiproc or handshake or linked or peer" ) 2>&1 | tail -8 | tee gpurun_out/r2z6_gpu_linked.log
exit 0
