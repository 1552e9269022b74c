#!/bin/bash
# GPU session Z (round 1, 2 GPUs): parity of the replay write-back skip and the 80-register moment variant,
# default bench at N=1 and N=2 (two processes linked through CUDA IPC across two devices).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multiproc.py -m gpu -q -x -k "porous or fuzz or staged or slab or ipc" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/z_n1.json 2>gpurun_out/z.err
python - <<'P'
import json
d=json.load(open("gpurun_out/z_n1.json"))
print("N=1", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), d["clocks"])
P
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
     bench.py --gpus 2 --steps 100 --warmup 10 2>gpurun_out/z_n2.err | grep '^{' > gpurun_out/z_n2.json
python - <<'P'
import json
try:
    d=json.load(open("gpurun_out/z_n2.json"))
    print("N=2", round(d["value"]), "MLUPS ms/step", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), "e2e", d["e2e"] and round(d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("N=2 FAILED", e); print(open("gpurun_out/z_n2.err").read()[-1500:])
P
