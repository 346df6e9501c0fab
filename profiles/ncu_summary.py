#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + per-region SASS execution counts / thread efficiency / stall samples.
usage: python profiles/ncu_summary.py <report.ncu-rep> [launch_index]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg"]
print("== raw metrics (launch %d) ==" % which)
for name in want:
    if name in hdr:
        i = hdr.index(name)
        print("%-75s %-14s %s" % (name, units[i], rows[2 + which][i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[which]]
end = hi[which + 1] - 1 if len(hi) > which + 1 else len(rows)
body = [r for r in rows[hi[which] + 1:end] if len(r) == len(h)]
ci = {n: i for i, n in enumerate(h)}
ti = sum(int(r[ci["Instructions Executed"]]) for r in body)
tt = sum(int(r[ci["Thread Instructions Executed"]]) for r in body)
print("== SASS: %d instructions, warp-instr executed %d, thread-instr %d, avg active threads %.2f ==" % (len(body), ti, tt, tt / max(ti, 1)))
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ci[n]]) for r in body) for n in stall_cols}
tot_s = sum(agg.values())
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(tot_s, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
seg = []
for k, r in enumerate(body):
    ie = int(r[ci["Instructions Executed"]])
    te = int(r[ci["Thread Instructions Executed"]])
    sm = int(r[ci["# Samples"]])
    op = r[ci["Source"]].strip()
    if seg and abs(seg[-1]["ie"] - ie) <= max(2, 0.03 * seg[-1]["ie"]):
        s = seg[-1]
        s["n"] += 1; s["thr"] += te; s["inst"] += ie; s["samples"] += sm; s["end"] = k
    else:
        seg.append(dict(start=k, end=k, ie=ie, n=1, thr=te, inst=ie, samples=sm, first=op))
print("== regions (runs of instructions with equal execution count; >0.5% of warp-instructions) ==")
for s in seg:
    if s["inst"] > 0.005 * ti:
        print("%4d-%4d n=%3d exec=%9d share=%5.1f%% avg_threads=%5.1f samples=%6d  %s" % (
            s["start"], s["end"], s["n"], s["ie"], 100 * s["inst"] / ti, s["thr"] / max(s["inst"], 1), s["samples"], s["first"][:48]))
