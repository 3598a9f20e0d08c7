"""Development aid: registers / spills of the tiled kernels from the ptxas -v logs."""
import re, sys, os
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "latticeurbanwind_b200", "lib")
for name in ("fast", "strict"):
    txt = open(os.path.join(lib, f"ptxas_{name}.log")).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
        k = m.group(1)
        t = re.search(r"TileCfgILi(\d)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)EEELj(\d+)ELb(\d)", k)
        if t:
            p, tx, ty, tz, s, c, feat, fast = t.groups()
            print(f"{name:6s} P={p} tile={tx}x{ty}x{tz} S={s} C={c} feat={feat:>2s}: regs={m.group(5):>3s} stack={m.group(2):>4s} spill_st={m.group(3):>4s} spill_ld={m.group(4):>4s}")
