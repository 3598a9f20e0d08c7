#!/bin/bash
mkdir -p gpurun_out
for ax in 0 1 2; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/halo_ax$ax.csv python dev/halo_times.py $ax > gpurun_out/halo_ax$ax.log 2>&1
tail -1 gpurun_out/halo_ax$ax.log
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/halo_ax$ax.csv")))
i=[k for k,r in enumerate(rows) if r and r[0]=="ID"][0]
agg=collections.defaultdict(list)
for r in rows[i+1:]:
    name=r[4].split("(")[0][-60:]
    agg[name].append(float(r[-1])/1e3)
for k,v in agg.items(): print("axis $ax", k, len(v), "launches, last us:", [round(x,1) for x in v[-4:]])
PY
done
