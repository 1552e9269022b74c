#!/bin/bash
# Multi-GPU session (round 1): weak scaling of the default workload and the 65536^2 channel on 8 GPUs.
N=${1:-8}
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() { # n workload extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2951$1 \
     bench.py --gpus $1 --steps 100 --warmup 10 --workload $2 $3 2>gpurun_out/scale_$1_$2.err | grep '^{' > gpurun_out/scale_$1_$2.json
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/scale_$1_$2.json"))
    print("N=$1 $2", round(d["value"]), "MLUPS ms/step", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "e2e", d["e2e"] and round(d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("N=$1 $2 FAILED", e); print(open("gpurun_out/scale_$1_$2.err").read()[-1500:])
P
}
if [ "$N" = "8" ]; then
  run 8 porous16384
  run 8 channel65536 --no-e2e
  run 4 porous16384 --no-e2e
else
  run $N porous16384
fi
