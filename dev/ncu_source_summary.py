"""Development aid: dynamic instruction mix and stall profile of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`.
usage: python dev/ncu_source_summary.py file.csv [kernel-index]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# the file holds one block per profiled launch: "Kernel Name" line, header line, instruction lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
b = blocks[which]
h = b["hdr"]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [(k, h.index(k)) for k in h if k.startswith("stall_") and "Not Issued" not in k]
ops, samples = collections.Counter(), collections.Counter()
tot = 0
stalls = collections.Counter()
for r in b["rows"]:
    src = r[iS].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    n = int(r[iE]); tot += n
    ops[op] += n; samples[op] += int(r[iN])
    for k, i in stall_cols:
        stalls[k] += int(r[i])
print(b["name"][:120])
print("warp instructions executed:", tot)
for op, n in ops.most_common(30):
    print(f"  {op:10s} {n:12d} {100*n/tot:5.1f}%   samples {samples[op]}")
ts = sum(stalls.values())
print("stall samples:", {k: f"{100*v/ts:.1f}%" for k, v in stalls.most_common(10)})
if len(sys.argv) > 3:  # top-N hottest instructions by samples
    top = sorted(b["rows"], key=lambda r: -int(r[iN]))[:int(sys.argv[3])]
    for r in top:
        print(r[iN], r[iE], r[iS].strip()[:100])
