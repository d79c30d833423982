"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep captured with
--import-source on: joins `ncu --page source --csv` (SASS order) with nvdisasm's line table of the same .so.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex for ncu -k> <mangled-name substring> [top N]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kregex, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
so = os.path.join(ROOT, "hyperqueue_b200", "libhqsched_b200.so")
_src_cache = {}


def src_line(path, line):
    """Source text of (file, line); the kernels live in hqsched.cu and the .cuh files it includes."""
    if path is None or line is None:
        return ""
    local = os.path.join(ROOT, "hyperqueue_b200", "csrc", os.path.basename(path))
    if local not in _src_cache:
        _src_cache[local] = open(local).read().split("\n") if os.path.exists(local) else []
    lines = _src_cache[local]
    return lines[line - 1].strip() if 0 < line <= len(lines) else ""

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
ins, cur = [], None
for l in dis[start + 1:]:
    if l.startswith("//--------------------- "):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kregex], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")][:len(ins)]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
by = defaultdict(lambda: [0, 0])
stalls = defaultdict(int)
for k, r in enumerate(data):
    by[ins[k][1]][0] += int(r[iI])
    by[ins[k][1]][1] += int(r[iS])
    for i in stall_cols:
        stalls[hdr[i]] += int(r[i] or 0)
tot_i = sum(v[0] for v in by.values())
tot_s = sum(v[1] for v in by.values())
print(f"SASS instructions {len(ins)}, warp-instructions executed {tot_i}, samples {tot_s}")
print("stalls:", sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
for l, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    where = f"{os.path.basename(l[0])}:{l[1]}" if l else "?"
    print(f"{where:>24} {v[0]:>9} {100.0 * v[0] / tot_i:5.1f}% {v[1]:>6}  {src_line(*l)[:100] if l else ''}")
