#!/bin/bash
# GPU session I of round 2 (1 GPU): compute-sanitizer over the kernels this round rewrote (ordered chain table, rank
# merge in the moment-storing launch, rank-based evict, staging ring, graph runs), and the small-lattice bench lines
# with the graphs captured in the warm-up.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for wl in cylinder512 cavity4096; do
  python bench.py --workload $wl --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/r2i_bench_${wl}_steps20.json 2>> gpurun_out/r2i_bench.err
done
python bench.py --workload cylinder512 --no-cpu-baseline > gpurun_out/r2i_bench_cylinder512.json 2>> gpurun_out/r2i_bench.err
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "chain_table_settles or paint_frames or create_state or (random_scripts and (size1-2 or size3-2 or size6-2 or size1-1-1)) or single_cell or color_maps" ) \
    2>&1 | tail -25 > gpurun_out/r2i_memcheck.log
tail -6 gpurun_out/r2i_memcheck.log
( time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "chain_table_settles and 1.0-2" ) 2>&1 | tail -15 > gpurun_out/r2i_racecheck.log
tail -6 gpurun_out/r2i_racecheck.log
tail -3 gpurun_out/r2i_bench.err
exit 0
