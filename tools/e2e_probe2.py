"""torchrun probe: per-step wall times of the sharded host-buffer loop (bench.py's e2e) for one workload"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
task, n, mode = sys.argv[1], int(sys.argv[2]), sys.argv[3]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{lr}")
g = bench.GpuRun(task, n, mode, "auto", rank, world, lr, 10, 3)
r = g.timed(flush)
K = 12
h_act = (torch.rand(K, n, g.A, generator=g.gen, device=g.dev) * 2 - 1).cpu().pin_memory()
h_out = torch.empty(n * world, g.O + 4).pin_memory()
torch.cuda.synchronize()
if world > 1: dist.barrier()
ts = []
for t in range(K):
    t0 = time.perf_counter()
    a = h_act[t].to(g.dev, non_blocking=True)
    g.env.step_packed(a, out=g.rec)
    torch.cuda.current_stream().synchronize(); t1 = time.perf_counter()
    if world > 1: dist.all_gather_into_tensor(g.full, g.rec)
    torch.cuda.current_stream().synchronize(); t2 = time.perf_counter()
    h_out.copy_(g.full if world > 1 else g.rec, non_blocking=True)
    torch.cuda.current_stream().synchronize(); t3 = time.perf_counter()
    ts.append((round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2), round((t3 - t2) * 1e3, 2)))
print(f"rank {rank}: timed {r['ms_per_step']:.2f} ms/step; per step (step, gather, d2h) ms:", ts, flush=True)
if world > 1: dist.destroy_process_group()
