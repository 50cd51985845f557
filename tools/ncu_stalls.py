#!/usr/bin/env python
"""Per-source-line stall breakdown of one ncu report (source page joined with nvdisasm line info).
usage: ncu_stalls.py <report.ncu-rep> <cubin> <kernel-substring> <stall column, e.g. stall_long_sb> [top]"""
import csv, os, re, subprocess, sys
from collections import defaultdict
rep, cubin, kname, col = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
sass = subprocess.run(["nvdisasm", "-c", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, active = [], None, False
for ln in sass:
    if ln.startswith("\t.section\t.text."):
        active = kname in ln; continue
    if ln.startswith("\t.section"): active = False
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln): lines.append((cur, ln.strip()))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, body = rows[hi], rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
agg, tot, insts = defaultdict(int), 0, defaultdict(lambda: defaultdict(int))
for k in range(min(len(body), len(lines))):
    v = int(body[k][ci[col]] or 0)
    agg[lines[k][0]] += v; tot += v
    op = re.sub(r"/\*[0-9a-f]+\*/\s*", "", lines[k][1]).split(";")[0][:60]
    insts[lines[k][0]][op] += v
print("total", col, tot)
src = {}
for l, v in sorted(agg.items(), key=lambda x: -x[1])[:top]:
    if l and l[0] not in src:
        try: src[l[0]] = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gym_lowcostrobot_b200", "csrc", l[0])).read().splitlines()
        except OSError: src[l[0]] = []
    txt = src[l[0]][l[1] - 1].strip()[:90] if l and len(src[l[0]]) >= l[1] else ""
    topop = max(insts[l].items(), key=lambda x: x[1])[0]
    print(f"{100*v/max(tot,1):6.2f}%  {l[0] if l else '?':16s}{l[1] if l else 0:5d}  {txt}   <- {topop}")
