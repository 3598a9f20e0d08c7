#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_W.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/halo2_ax0.csv python dev/halo_times.py 0 > gpurun_out/halo2_ax0.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/halo2_ax0.csv")))
i=[k for k,r in enumerate(rows) if r and r[0]=="ID"][0]
agg=collections.defaultdict(list)
for r in rows[i+1:]:
    agg[r[4].split("(")[0][-60:]].append(float(r[-1])/1e3)
for k,v in agg.items(): print("axis 0", k, len(v), "launches, last us:", [round(x,1) for x in v[-4:]])
PY
