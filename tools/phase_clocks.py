#!/usr/bin/env python
"""Per-env phase timing of the lockstep kernel (debug hook lcr_debug_phase_clocks): where do the warps spend their cycles?
usage: phase_clocks.py [task] [n_envs] [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gym_lowcostrobot_b200 as glr
task, n, T = (sys.argv[1] if len(sys.argv) > 1 else "ReachCube-v0"), int(sys.argv[2]) if len(sys.argv) > 2 else 4096, int(sys.argv[3]) if len(sys.argv) > 3 else 25
env = glr.make(task, num_envs=n, autoreset=True, exec_mode="lockstep")
env.reset(seed=0)
g = torch.Generator(device="cuda").manual_seed(1234)
buf = torch.zeros(n, 10, dtype=torch.int64, device="cuda")
for t in range(T):
    if t == T - 1:
        env._L.lcr_debug_phase_clocks(env._h, C.c_void_p(buf.data_ptr()))
    env.step(torch.rand(n, env.action_dim, generator=g, device="cuda") * 2 - 1)
torch.cuda.synchronize()
c = buf.cpu().numpy().astype(np.float64)
names = ["begin/end", "wait top", "dyn+broad", "wait pre-pool", "jobs", "wait post-pool", "rows+smooth", "wait pre-solve", "solve", "integrate"]
tot = c.sum(1)
print(f"{task} n={n}: per-env total cycles mean {tot.mean():.0f} max {tot.max():.0f}  (kernel ~ {tot.max()/1.965e6:.2f} ms per CTA round)")
for k, nm in enumerate(names):
    print(f"  {nm:16s} mean {c[:, k].mean():10.0f} ({100 * c[:, k].mean() / tot.mean():5.1f}%)  p50 {np.median(c[:, k]):10.0f}  p99 {np.percentile(c[:, k], 99):10.0f}  max {c[:, k].max():10.0f}")
