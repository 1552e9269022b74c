#!/bin/bash
# GPU session Y (round 1): streaming-store hint A/B (porous, channel), block shapes again under the new issue
# order, and an ncu --set full capture of the reworked staged kernel (where do the stalls sit now?).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fuzz or staged or random_scripts" 2>&1 | tail -3
show() { python - "$1" <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "launch ms", round(d["roofline"]["avg_launch_ms"],4), d["clocks"])
P
}
for wl in porous16384 channel16384; do
  for ss in 0 1 0 1; do
    timeout 300 python bench.py --workload $wl --stream-stores $ss --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
        > gpurun_out/y_${wl}_ss${ss}.json 2>>gpurun_out/y.err
    show gpurun_out/y_${wl}_ss${ss}.json
  done
done
for rows in 2 8; do
  timeout 300 python bench.py --workload porous16384 --block-rows $rows --steps 100 --warmup 10 --no-cpu-baseline --no-e2e \
      > gpurun_out/y_porous16384_rows${rows}.json 2>>gpurun_out/y.err
  show gpurun_out/y_porous16384_rows${rows}.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_vec4_kernel -s 20 -c 1 -o gpurun_out/prof_y_porous16384 \
   python bench.py --workload porous16384 --steps 12 --warmup 12 --no-cpu-baseline --no-e2e > gpurun_out/ncu_y_porous.log 2>&1
tail -1 gpurun_out/ncu_y_porous.log | cut -c1-200
