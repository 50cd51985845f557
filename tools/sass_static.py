#!/usr/bin/env python
"""Static SASS instruction count per source function for one kernel (nvdisasm -gi line info).
usage: sass_static.py <cubin> <kernel-mangled-substring>"""
import os, re, subprocess, sys
from collections import defaultdict
cubin, kname = sys.argv[1:3]
sass = subprocess.run(["nvdisasm", "-c", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
srcdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gym_lowcostrobot_b200", "csrc")
func_of = {}
for fn in os.listdir(srcdir):
    if not fn.endswith((".cuh", ".cu")):
        continue
    curf = "?"
    for i, l in enumerate(open(os.path.join(srcdir, fn)).read().splitlines(), 1):
        m2 = re.match(r"^(?:template <typename T> )?(?:__device__ __noinline__|DI|__global__)\s+.*?(\w+)\((?!.*;\s*$)", l) or re.match(r"^\w[\w<>:, \*&]*\s+(\w+)\(.*\)\s*\{?$", l)
        if m2 and not l.startswith(" "):
            curf = m2.group(1)
        func_of[(fn, i)] = curf
active, cur, cnt, tot = False, None, defaultdict(int), 0
for ln in sass:
    if ln.startswith("\t.section\t.text."):
        active = kname in ln
        continue
    if ln.startswith("\t.section"):
        active = False
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln):
        cnt[func_of.get(cur, "?")] += 1
        tot += 1
print("total", tot, "instrs", tot * 16 / 1024, "KB")
for f, c in sorted(cnt.items(), key=lambda x: -x[1]):
    print(f"{f:28s} {c:6d} {c*16/1024:7.1f} KB")
