#!/bin/bash
# Strong scaling of BASELINE.json configs[3] (32768^2 cylinder wake) at N GPUs.
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N \
   bench.py --gpus $N --steps 100 --warmup 10 --workload cylinder32768 --no-e2e 2>gpurun_out/strong_$N.err | grep '^{' > gpurun_out/strong_$N.json
python - <<P
import json
try:
    d=json.load(open("gpurun_out/strong_$N.json"))
    print("N=$N cylinder32768", round(d["value"]), "MLUPS ms/step", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), d["scaling"], d["clocks"])
except Exception as e:
    print("N=$N FAILED", e); print(open("gpurun_out/strong_$N.err").read()[-1500:])
P
