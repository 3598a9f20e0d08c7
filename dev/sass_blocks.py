"""Development aid: basic blocks of one kernel's SASS (cuobjdump -sass), with instruction counts and opcode mix per block.
usage: python dev/sass_blocks.py file.cubin kernel-name-substring [--list]"""
import collections, re, subprocess, sys
cubin, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = next(f for f in funcs if key in f.split("\n")[0])
ins = []
for line in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
targets = set()
for addr, s in ins:
    for m in re.finditer(r"0x([0-9a-f]+)", s):
        if re.search(r"\b(BRA|BSSY|CALL|BRX|JMP)\b", s):
            targets.add(int(m.group(1), 16))
blocks, cur = [], []
for addr, s in ins:
    if addr in targets and cur:
        blocks.append(cur); cur = []
    cur.append((addr, s))
    if re.search(r"\b(BRA|EXIT|RET|BRX|JMP)\b", s) :
        blocks.append(cur); cur = []
if cur: blocks.append(cur)
def op(s):
    t = s.split()
    o = t[1] if t[0].startswith("@") else t[0]
    return o.split(".")[0]
print(len(ins), "instructions,", len(blocks), "blocks")
for b in blocks:
    c = collections.Counter(op(s) for _, s in b)
    last = b[-1][1]
    print(f"{b[0][0]:06x} n={len(b):4d}  end: {last[:60]:60s} " + " ".join(f"{k}:{v}" for k, v in c.most_common(8)))
    if "--list" in sys.argv:
        for a, s in b: print(f"      {a:06x} {s}")
