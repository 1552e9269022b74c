#!/bin/bash
# GPU session A of round 2 (prepared at the end of round 1, when the GPU budget was spent): first the tests that
# were added without a GPU at hand — the executed-WGSL digests of configs[0] (10,000 steps), of configs[1] at
# full size and of the large scripts, single domain and slab groups — then the whole GPU suite, the default
# bench line with its launch list, and one full ncu capture of the step kernel.
#   gpurun --timeout 2400 -- bash profiles/run_round2_a.sh
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_wgsl_pin.py -m gpu -q -x -k "config1 or large_lattices or slab_group" ) 2>&1 | tail -15 | tee gpurun_out/r2a_wgsl_new.log
( time timeout 1500 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -15 | tee gpurun_out/r2a_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2a_smoke.log
python bench.py > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
tail -1 gpurun_out/r2a_bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_default.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_vec4_kernel -s 20 -c 1 -o gpurun_out/r2a_step_vec4 \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2a_ncu_full.log 2>&1
ncu -i gpurun_out/r2a_step_vec4.ncu-rep --page raw --csv > gpurun_out/r2a_step_vec4_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r2a_step_vec4_raw.csv > gpurun_out/r2a_step_vec4_summary.txt 2>&1 || true
rm -f gpurun_out/r2a_step_vec4.ncu-rep
