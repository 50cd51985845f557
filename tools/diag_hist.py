#!/usr/bin/env python
"""Distribution of constraint counts and Newton iterations over a batch (lcr_get_diag), after T random steps.
usage: diag_hist.py [task-id] [n_envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gym_lowcostrobot_b200 as glr
task, n, T = (sys.argv[1] if len(sys.argv) > 1 else "ReachCube-v0"), int(sys.argv[2]) if len(sys.argv) > 2 else 4096, int(sys.argv[3]) if len(sys.argv) > 3 else 25
env = glr.make(task, num_envs=n, autoreset=True, exec_mode="lockstep")
env.reset(seed=0)
g = torch.Generator(device="cuda").manual_seed(1234)
it_hist = np.zeros(64, np.int64)
for t in range(T):
    env.step(torch.rand(n, env.action_dim, generator=g, device="cuda") * 2 - 1)
    if t >= T - 5:
        d = env.diagnostics()
        d = {k: v.cpu().numpy() for k, v in d.items()} if isinstance(d, dict) else d.cpu().numpy()
        arr = d if not isinstance(d, dict) else np.stack([d[k] for k in d], 1)
        it_hist += np.bincount(np.minimum(arr[:, 2], 63), minlength=64)
print("columns: ncon nefc niter(last substep) max_nefc overflow nan_resets")
print("last-substep Newton iterations histogram (5 steps x n envs):", {i: int(c) for i, c in enumerate(it_hist) if c})
for lo, hi in [(0, 17), (17, 25), (25, 41), (41, 65), (65, 200)]:
    msk = (arr[:, 1] >= lo) & (arr[:, 1] < hi)
    if msk.any():
        print(f"nefc in [{lo},{hi}): {msk.sum():5d} envs, niter mean {arr[msk, 2].mean():.2f} max {arr[msk, 2].max()}, ncon mean {arr[msk, 0].mean():.1f}")
