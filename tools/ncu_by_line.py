#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page by CUDA source line / function.

usage: ncu_by_line.py <report.ncu-rep> <cubin> <kernel-mangled-name-substring> [top]
Joins `ncu --page source --csv` (per SASS instruction: samples, executed) with `nvdisasm -gi` line
info by instruction order inside the kernel's .text section.
"""
import csv
import os
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["nvdisasm", "-c", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, active, inl = [], None, False, None
for ln in sass:
    if ln.startswith("\t.section\t.text."):
        active = kname in ln
        continue
    if ln.startswith("\t.section"):
        active = False
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = rows[hi + 1:]
ci = {h: i for i, h in enumerate(hdr)}
print("sass rows", len(body), "disasm instrs", len(lines))
srcs, func_of = {}, {}
for fn in ("lcr_kernels.cuh", "lcr_convex.cuh", "lcr_device.cuh"):
    src = open("/root/repo/gym_lowcostrobot_b200/csrc/" + fn).read().splitlines()
    srcs[fn] = src
    curf = "?"
    for i, l in enumerate(src, 1):
        m2 = re.match(r"^(?:template <typename T> )?(?:__device__ __noinline__|DI|__global__)\s+.*?(\w+)\((?!.*;\s*$)", l) or re.match(r"^\w[\w<>:, \*&]*\s+(\w+)\(.*\)\s*\{?$", l)
        if m2 and not l.startswith(" "):
            curf = m2.group(1)
        func_of[(fn, i)] = curf
by_line = defaultdict(lambda: [0, 0])
tot_s = tot_e = 0
n = min(len(body), len(lines))
for k in range(n):
    r = body[k]
    s = int(r[ci["# Samples"]] or 0)
    e = int(r[ci["Instructions Executed"]] or 0)
    by_line[lines[k]][0] += s
    by_line[lines[k]][1] += e
    tot_s += s
    tot_e += e
by_func = defaultdict(lambda: [0, 0])
for l, (s, e) in by_line.items():
    by_func[func_of.get(l, "?")][0] += s
    by_func[func_of.get(l, "?")][1] += e
print(f"total samples {tot_s} executed {tot_e}")
print("== by function (samples%, executed%)")
for f, (s, e) in sorted(by_func.items(), key=lambda x: -x[1][0])[:25]:
    print(f"{f:28s} {100*s/tot_s:6.2f}% {100*e/tot_e:6.2f}%")
print("== by line")
for l, (s, e) in sorted(by_line.items(), key=lambda x: -x[1][0])[:top]:
    txt = srcs.get(l[0], [""] * 100000)[l[1] - 1].strip()[:100] if l else ""
    print(f"{l[0][4:12] if l else '':8s}{l[1] if l else 0:5d} {100*s/tot_s:6.2f}% {100*e/tot_e:6.2f}%  {txt}")
