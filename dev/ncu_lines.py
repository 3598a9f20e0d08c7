"""Development aid: per-CUDA-source-line executed instructions / stall samples from `ncu --page source --csv --print-source cuda,sass`.
usage: python dev/ncu_lines.py file.csv [top-N]"""
import csv, sys, collections
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
fname = None
hdr = None
seen_first_launch = set()
launch = 0
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == "Kernel Name": launch += 1; continue
    if r[0] in ("File Path", "File Name"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or launch > 1: continue
    if r[0] != "" and r[0].isdigit():
        iE, iN = hdr.index("Instructions Executed"), hdr.index("# Samples")
        key = (fname, int(r[0]))
        e = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        e[0] += int(r[iE]) if r[iE].isdigit() else 0; e[1] += int(r[iN]) if r[iN].isdigit() else 0
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp instr", tot, "samples", ts)
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{l:4d} {v[0]:11d} {100*v[0]/tot:5.1f}% smp {100*v[1]/max(ts,1):5.1f}%  {v[2]}")
