#!/bin/bash
# GPU session AI (round 1): after the fill_equilibrium fix — slab groups, staged / packed flavours, KATs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "slab_group_on_one or staged or packed or closed_box or conserves or translates or color_maps" > gpurun_out/ai.log 2>&1
tail -4 gpurun_out/ai.log
