#!/usr/bin/env python
"""Turn the scratch outputs of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: summarize_profiles.py <tag>    e.g. r01a"""
import collections, csv, glob, json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
os.makedirs(P, exist_ok=True)

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "second": 1e9}.get(row["Metric Unit"], 1)
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# {os.path.basename(path)}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)",
           f"# total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches", "kernel,launches,total_ms,avg_us,share"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k},{v[0]},{v[1]/1e6:.3f},{v[1]/v[0]/1e3:.1f},{v[1]/tot:.4f}")
    return "\n".join(out) + "\n"

for f in sorted(glob.glob(os.path.join(G, "launches_*.csv"))):
    open(os.path.join(P, f"{tag}_{os.path.basename(f)[:-4]}_summary.csv"), "w").write(launches(f))
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum ", "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "launch__registers_per_thread ", "launch__occupancy_limit", "launch__grid_size", "launch__block_size", "launch__waves",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct", "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_per_inst_issued",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct", "smsp__sass_inst_executed_op_local", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_fma", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for rep in sorted(glob.glob(os.path.join(G, "prof_*.ncu-rep"))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"# {os.path.basename(rep)}: ncu --set full --clock-control none --import-source on (selected raw metrics, one column per captured launch)"]
    for ci, (h, u) in enumerate(zip(hdr, units)):
        if any(k.strip() in h for k in KEYS) or h in ("Kernel Name", "ID"):
            out.append(",".join([h, u] + [r[ci].split("(")[0] for r in rows[2:]]))
    open(os.path.join(P, f"{tag}_{os.path.basename(rep)[:-8]}_raw.csv"), "w").write("\n".join(out) + "\n")
for f in sorted(glob.glob(os.path.join(G, "bench_*.json"))):
    s = open(f).read().strip()
    if s:
        open(os.path.join(P, f"{tag}_{os.path.basename(f)}"), "w").write(s + "\n")
for f in ("pytest_gpu.log", "smoke.log"):
    if os.path.exists(os.path.join(G, f)):
        open(os.path.join(P, f"{tag}_{f}"), "w").write(open(os.path.join(G, f)).read())
print("wrote", sorted(x for x in os.listdir(P) if x.startswith(tag)))
