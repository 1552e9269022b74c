#!/bin/bash
# GPU session B (round 1): parity with the barrier-chain table, porous bench dense vs lazy.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 600 python bench.py --workload porous16384 --no-cpu-baseline > gpurun_out/b_porous_lazy.json 2> gpurun_out/b_porous_lazy.err
cut -c1-1400 gpurun_out/b_porous_lazy.json; tail -3 gpurun_out/b_porous_lazy.err
timeout 600 python bench.py --workload channel16384 --no-cpu-baseline --no-e2e > gpurun_out/b_channel.json 2>&1
cut -c1-1000 gpurun_out/b_channel.json
