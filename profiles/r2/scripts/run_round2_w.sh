#!/bin/bash
# GPU session W (1 GPU): ncu evidence of the FINAL library: launch list of the default bench command, one --set full capture of
# the default step kernel (traffic.json) and of the small-lattice graph replay.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2w_launches_default.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2w_launches_cylinder512.csv \
    python bench.py --workload cylinder512 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_launches_cylinder512.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"step_vec4_kernel<\(bool\)0" -s 20 -c 1 -o gpurun_out/r2w_step_vec4 \
    python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2w_ncu_full.log 2>&1
python profiles/ncu_summary.py gpurun_out/r2w_step_vec4.ncu-rep > gpurun_out/r2w_ncu_step_vec4.txt 2>&1 || true
rm -f gpurun_out/r2w_step_vec4.ncu-rep
tail -3 gpurun_out/r2w_ncu_full.log
exit 0
