"""GPU experiment: phased-chain tunables (env vars read by lcr_create) in the bench's stationary window.

  python tools/ph_sweep.py push joint 16384 "LCR_GRAPH=0" "LCR_GRAPH=1 LCR_GROUPS=4" ...

Every configuration runs the same seeds and actions; prints ms/step (device), the host time spent enqueueing a step,
and a checksum of the last output record (all configurations must print the same one: the tunables never change results).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

task, mode, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
K = int(os.environ.get("SWEEP_STEPS", 12))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
base = dict(os.environ)
for spec in sys.argv[4:]:
    os.environ.clear()
    os.environ.update(base)
    em = "auto"
    for kv in spec.split():
        k, v = kv.split("=")
        if k == "EXEC":
            em = v
        else:
            os.environ[k] = v
    g = bench.GpuRun(task, n, mode, em, 0, 1, 0, K, 3)
    r = g.timed(flush)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(K):
        g.env.step_packed(g.actions[3 + t], out=g.rec)
    host = (time.perf_counter() - t0) / K * 1e3
    torch.cuda.synchronize()
    chk = float(g.rec.double().sum().item())
    d = g.env.diagnostics()
    line = f"{spec:44s} {r['ms_per_step']:8.3f} ms/step {r['value']:10.0f} env-steps/s  host {host:6.3f} ms  launches/step {r['launches'] / K:6.1f}  chk {chk:.9e}"
    if os.environ.get("SWEEP_HIST"):
        it = torch.bincount(d["niter"].clamp(max=39) // 4, minlength=10).tolist()
        ne = torch.bincount(d["max_nefc"].clamp(max=199) // 20, minlength=10).tolist()
        line += f"\n    niter/4 hist {it}  max_nefc/20 hist {ne}  max niter {int(d['niter'].max())}"
    print(line, flush=True)
    g.close()
