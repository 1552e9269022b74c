#!/bin/bash
# GPU session Q (round 1, re-entry): full GPU suite + smoke + default bench + reference arm at HEAD
# (validates the stale-halo fix and the API fuzz tests of the last commit on a fresh box).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) 2>&1 | tail -16
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/q_default.json 2>gpurun_out/q.err; tail -2 gpurun_out/q.err
cut -c1-1800 gpurun_out/q_default.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 | cut -c1-400
