#!/usr/bin/env python
"""Per source line: executed instructions, stall samples and the dominant stall reasons.

Joins the SASS table of an .ncu-rep (ncu --page source --print-source sass: no line numbers)
with `nvdisasm -g` of the object file the kernel came from (compiled with -lineinfo) by
instruction offset.

usage: python tools/ncu_lines.py prof.ncu-rep build/mcmcn_sets_tc.o 'sweep_tc_kernelILi1E' [top]
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

rep, obj, pattern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}
cur, inside = None, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        inside = pattern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
d = [r for r in rows[2:] if len(r) == len(h) and r[0].startswith("0x")]
ci, si, sc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
stalls = [(i, n.replace("stall_", "")) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
base = int(d[0][0], 16)
agg = defaultdict(lambda: [0, 0, Counter(), Counter()])
tot_i = tot_s = 0
for r in d:
    key = line_of.get(int(r[0], 16) - base)
    e = agg[key]
    e[0] += int(r[ci]); e[1] += int(r[si])
    tot_i += int(r[ci]); tot_s += int(r[si])
    for i, n in stalls:
        if r[i] and int(r[i]):
            e[2][n] += int(r[i])
    op = r[sc].strip().split()
    op = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    e[3][op] += int(r[ci])
print("total executed %d, samples %d" % (tot_i, tot_s))
src_cache = {}
def src(key):
    if not key:
        return "?"
    f, n = key
    if f not in src_cache:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)) + glob.glob(os.path.join("include", f))
        src_cache[f] = open(cands[0]).read().splitlines() if cands else []
    L = src_cache[f]
    return L[n - 1].strip()[:70] if 0 < n <= len(L) else ""
print("%-22s %7s %7s  %-38s %s" % ("line", "exec%", "smpl%", "top stalls", "source"))
for key, e in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = " ".join("%s:%d" % (n, 100 * v // max(e[1], 1)) for n, v in e[2].most_common(3))
    print("%-22s %6.1f%% %6.1f%%  %-38s %s" % ("%s:%d" % key if key else "?", 100.0 * e[0] / tot_i, 100.0 * e[1] / tot_s, st, src(key)))
