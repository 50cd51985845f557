"""GPU experiment: kernel timeline of the phased chain with the concurrency between the env groups preserved (CUPTI through
torch.profiler; ncu serialises the launches).  Writes one compact JSON list [name, stream, start_us, dur_us] per profiled step.

  python tools/ph_timeline.py push joint 16384 gpurun_out/timeline_push16384.json [steps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

task, mode, n, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
g = bench.GpuRun(task, n, mode, "auto", 0, 1, 0, steps, 3)
for t in range(3):
    g.env.step_packed(g.actions[t], out=g.rec)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for t in range(steps):
        g.env.step_packed(g.actions[3 + t], out=g.rec)
    torch.cuda.synchronize()
tmp = out + ".trace.json"
prof.export_chrome_trace(tmp)
ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") == "kernel"]
t0 = min(e["ts"] for e in ev)
rows = sorted([[e["name"].split("<")[0].split("(")[0].replace("void lcr::", "") + ("[BIG]" if ", 17" in e["name"] or "(int)17" in e["name"] or "<float, 17" in e["name"] else ""),
                e["args"].get("stream", -1), round(e["ts"] - t0, 3), round(e["dur"], 3)] for e in ev], key=lambda r: r[2])
json.dump(rows, open(out, "w"))
os.remove(tmp)
print(len(rows), "kernels;", "span", round(max(r[2] + r[3] for r in rows) / 1e3, 3), "ms for", steps, "steps")
