"""Summarise one `ncu --set full` capture (raw + source pages exported as csv) into profiles/<tag>.md and profiles/traffic_<workload>.json.
usage: python dev/profile_summary.py <tag> <workload> <raw.csv> <src.csv> <cells_per_launch> <alg_bytes_per_cell>"""
import collections
import csv
import json
import sys

tag, workload, raw_csv, src_csv, cells, balg = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
rows = list(csv.reader(open(raw_csv)))
m = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
def val(k):
    return float(m[k][1].replace(",", ""))
def unit_scale(k):
    u = m[k][0]
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
t = val("gpu__time_duration.sum") * unit_scale("gpu__time_duration.sum")
rd = val("dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum")
wr = val("dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum")
json.dump({"workload": workload, "kernel": m["Kernel Name"][1], "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "algorithmic_bytes_per_launch": cells * balg, "gpu_time_s_under_ncu": t, "source": f"profiles/{tag}.md (ncu --set full --clock-control none)"},
          open(f"profiles/traffic_{workload}.json", "w"), indent=1)
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg"]
out = [f"# {tag}: ncu --set full --clock-control none, one launch of the step kernel ({workload})", "",
       f"kernel: `{m['Kernel Name'][1]}`", "",
       f"* duration under ncu (cold cache, serialised): {t*1e3:.3f} ms", f"* DRAM traffic per launch: {(rd+wr)/1e9:.3f} GB (read {rd/1e9:.3f} + write {wr/1e9:.3f})",
       f"* algorithmic bytes per launch: {cells*balg/1e9:.3f} GB ({cells} cells x {balg} B) -> traffic / algorithmic = {(rd+wr)/(cells*balg):.3f}",
       f"* DRAM throughput under ncu: {(rd+wr)/t/1e9:.0f} GB/s; algorithmic: {cells*balg/t/1e9:.0f} GB/s", "", "| metric | unit | value |", "|---|---|---|"]
for k in keys:
    if k in m:
        out.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
out += ["", "Warp stall reasons (average warps stalled per issued instruction):", "", "| reason | value |", "|---|---|"]
for h in rows[0]:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        v = val(h)
        if v >= 0.05:
            out.append(f"| {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {v:.3f} |")
# instruction mix from the SASS page
srows = list(csv.reader(open(src_csv)))
h = srows[1]
iS, iE = h.index("Source"), h.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for r in srows[2:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    tk = r[iS].split()
    op = (tk[1] if tk[0].startswith("@") else tk[0]).split(".")[0]
    ops[op] += int(r[iE]); tot += int(r[iE])
out += ["", f"Dynamic SASS mix ({tot} warp instructions, {tot/(cells/64):.0f} per 64 cells):", "", "| opcode | warp instructions | share |", "|---|---|---|"]
for op, n in ops.most_common(24):
    out.append(f"| {op} | {n} | {100*n/tot:.1f} % |")
tma = {op: n for op, n in ops.items() if op.startswith("UTMA") or op in ("SYNCS", "UBLKCP", "UTMALDG", "UTMASTG")}
out += ["", "TMA / mbarrier opcodes present: " + ", ".join(f"{k} x{v}" for k, v in sorted(tma.items()))]
open(f"profiles/{tag}.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
