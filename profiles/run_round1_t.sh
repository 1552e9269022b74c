#!/bin/bash
# GPU session T (round 1): (1) CUDA path vs the golden vectors interpreted from the reference's WGSL text;
# (2) cp.async-staged bounce-back (dense flavour 2) — parity, then same-box A/B against flavour 1.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wgsl_pin.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "staged or fuzz or random_scripts" 2>&1 | tail -3
for dn in 1 2 1 2; do
  timeout 300 python bench.py --workload porous16384 --dense $dn --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/t_porous16384_dense${dn}.json 2>>gpurun_out/t.err
  python - gpurun_out/t_porous16384_dense${dn}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
done
for rows in 2 8; do
  timeout 300 python bench.py --workload porous16384 --dense 2 --block-rows $rows --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/t_porous16384_dense2_rows${rows}.json 2>>gpurun_out/t.err
  python - gpurun_out/t_porous16384_dense2_rows${rows}.json <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
done
for dn in 1 2; do
  timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct \
     --clock-control none -k regex:step_vec4 -s 12 -c 2 --csv --log-file gpurun_out/t_ncu_porous_dense${dn}.csv \
     python bench.py --workload porous16384 --dense $dn --steps 6 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  tail -14 gpurun_out/t_ncu_porous_dense${dn}.csv | cut -d, -f5,13-
done
