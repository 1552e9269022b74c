#!/bin/bash
# GPU session P (round 1): compute-sanitizer memcheck + racecheck on small lattices (all kernels, slabs, paints).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_scripts and (size1 or size3 or size5 or size8) or slab_group or color_maps" 2>&1 | tail -4
echo "memcheck exit: $?"; tail -5 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "random_scripts and size1 and 3" 2>&1 | tail -4
echo "racecheck exit: $?"; tail -5 gpurun_out/racecheck.log
