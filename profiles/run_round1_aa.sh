#!/bin/bash
# GPU session AA (round 1): vectorised curl kernel, 80-register moment variant, replay write-back skip —
# full GPU suite, default bench, launch list of the e2e frames.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/aa_default.json 2>gpurun_out/aa.err; tail -2 gpurun_out/aa.err
python - <<'P'
import json
d=json.load(open("gpurun_out/aa_default.json"))
print("default", round(d["value"]), "MLUPS frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), d["clocks"], "cpu", round(d["cpu_baseline"]["value"]))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_aa_default.csv \
   python bench.py --steps 30 --warmup 15 --no-cpu-baseline > gpurun_out/ncu_launches_aa.log 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_aa_default.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; ki=H.index("Kernel Name"); vi=H.index("Metric Value")
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hdr+2:]:
    if len(r)>vi:
        k=r[ki].split('(')[0][:56]; tot[k]+=float(r[vi].replace(',','')); cnt[k]+=1
T=sum(tot.values())
for k,v in tot.most_common(9): print(f"{v/1e6:9.2f} ms {100*v/T:5.1f}%  n={cnt[k]:3d} avg {v/cnt[k]/1e6:.3f} ms  {k}")
P
