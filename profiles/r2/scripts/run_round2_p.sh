#!/bin/bash
# GPU session P (2 GPUs, ~1 minute): indirect NVLink evidence — peer copy bandwidth between two GPUs of the box.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python profiles/p2p_bandwidth.py > gpurun_out/r2p_p2p_bandwidth.json 2> gpurun_out/r2p.err
cat gpurun_out/r2p_p2p_bandwidth.json
nvidia-smi nvlink --status -i 0 > gpurun_out/r2p_nvlink_status.txt 2>&1
head -25 gpurun_out/r2p_nvlink_status.txt
tail -3 gpurun_out/r2p.err
exit 0
