#!/bin/bash
# GPU session AC (round 1): the new size tests (wide rows, 32768^2 offset widths) and the full suite timing.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide or 32768 or tall" --durations=5 ) 2>&1 | tail -14
free -g | head -2
