#!/bin/bash
# GPU session AD (round 1): bench lines of the other BASELINE.json configurations with the final kernels
# (configs[0] 512x256 cylinder, configs[1] 4096^2 closed box, the empty 16384^2 channel).
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - "$1" <<'P'
import json,sys
d=json.load(open(sys.argv[1]))
e=d.get("e2e")
print(sys.argv[1], round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "ms/step", round(d["ms_per_step"],5), "e2e", e and round(e["value"]), d["clocks"])
P
}
timeout 300 python bench.py --workload cylinder512 --steps 20000 --warmup 200 --no-cpu-baseline > gpurun_out/ad_cylinder512.json 2>gpurun_out/ad.err; show gpurun_out/ad_cylinder512.json
timeout 300 python bench.py --workload cavity4096 --steps 1000 --warmup 50 --no-cpu-baseline > gpurun_out/ad_cavity4096.json 2>>gpurun_out/ad.err; show gpurun_out/ad_cavity4096.json
timeout 300 python bench.py --workload channel16384 --steps 150 --warmup 15 --no-cpu-baseline > gpurun_out/ad_channel16384.json 2>>gpurun_out/ad.err; show gpurun_out/ad_channel16384.json
