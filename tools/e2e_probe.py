import sys, time; sys.path.insert(0, '/root/repo')
import torch, bench
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for task, n, mode in (("pick_place", 8192, "ee"), ("push", 16384, "joint")):
    g = bench.GpuRun(task, n, mode, "auto", 0, 1, 0, 15, 3)
    r = g.timed(flush)
    def loop(sync, copies, K=15):
        h_act = torch.empty(K, n, g.A).pin_memory(); h_act.copy_(g.actions[3:3+K].cpu())
        h_out = torch.empty(n, g.O + 4).pin_memory()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in range(K):
            a = h_act[t].to(g.dev, non_blocking=True) if copies else g.actions[3 + t]
            out = g.step(a)
            if copies: h_out.copy_(out, non_blocking=True)
            if sync: torch.cuda.current_stream().synchronize()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / K * 1e3
    print(task, "timed", round(r["ms_per_step"], 2), "| nosync", round(loop(False, False), 2), "| sync", round(loop(True, False), 2), "| sync+copies", round(loop(True, True), 2), "| again nosync", round(loop(False, False), 2), "e2e()", round(n / g.e2e()["value"] * 1e3, 2))
    g.close()
