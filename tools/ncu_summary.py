#!/usr/bin/env python
"""Summarise an .ncu-rep of the step kernel: headline metrics, stall mix, executed-instruction
classes by execution count (hot loop vs per-sweep vs per-group code).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, subprocess, sys, io
from collections import Counter, defaultdict
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print("%-75s %-12s %s" % (h, units[i], [r[i][:60] for r in data]))
print("-- stall reasons (warps per issue)")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = float(data[0][i])
        if v > 0.05:
            print("   %-28s %.2f" % (h.split("stalled_")[1].replace("_per_issue_active.ratio", ""), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
d = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
ci, si, sc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
tot = sum(int(r[ci]) for r in d); tots = sum(int(r[si]) for r in d)
cls = defaultdict(lambda: [0, 0, 0, Counter()])
maxc = max(int(r[ci]) for r in d)
for r in d:
    c = int(r[ci])
    key = "hot loop" if c > maxc * 0.5 else c
    ins = r[sc].strip().split()
    o = (ins[1] if ins[0].startswith("@") else ins[0]).split(".")[0]
    e = cls[key]; e[0] += 1; e[1] += c; e[2] += int(r[si]); e[3][o] += 1
print("-- SASS instructions: %d, executed %d, samples %d" % (len(d), tot, tots))
for k, e in sorted(cls.items(), key=lambda kv: -kv[1][1])[:6]:
    print("   exec-count %-9s n=%-5d executed %5.1f%%  samples %5.1f%%  %s" % (k, e[0], 100 * e[1] / tot, 100 * e[2] / tots, e[3].most_common(6)))
names = [(i, h[i]) for i in range(len(h)) if h[i].startswith("stall_") and "Not Issued" not in h[i]]
for label, sel in (("hot loop", lambda c: c > maxc * 0.5), ("rest", lambda c: c <= maxc * 0.5)):
    agg = Counter()
    for r in d:
        if sel(int(r[ci])):
            for j, nm in names:
                agg[nm] += int(r[j] or 0)
    t = sum(agg.values()) or 1
    print("   stalls in %-8s %s" % (label, [(k.replace("stall_", ""), round(100 * v / t, 1)) for k, v in agg.most_common(7)]))
