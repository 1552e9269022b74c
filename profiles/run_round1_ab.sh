#!/bin/bash
# GPU session AB (round 1): compute-sanitizer memcheck + racecheck on the reworked kernels (cp.async staging in
# shared memory, 2-D/3-D grid, vectorised curl, replay write-back skip) on small lattices.
# Single-process slab groups are left out: the sanitizer serialises kernels across streams, so a slab's spinning
# wait kernel can never see the sibling slab's signal (the 20 s timeout fires; first attempt of this session).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest tests/test_gpu_parity.py tests/test_wgsl_pin.py -m gpu -q -x -k "staged or packed or (random_scripts and (size1 or size3 or size5 or size8)) or colour or color_maps or edges or presets or porous_33 or (fuzz and not slab)" 2>&1 | tail -4
echo "memcheck exit: $?"; tail -4 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(staged and size1 and (4-2 or 16-2 or 1-3)) or (random_scripts and size1 and 2-)" 2>&1 | tail -4
echo "racecheck exit: $?"; tail -4 gpurun_out/racecheck.log
