"""Tiny run for compute-sanitizer over the round-2 code paths: phased chain replayed as a graph, envs that outgrow the fast caps
mid-step (migration lists + BIG resume passes), the predicted-BIG pass, and the recorder kernel.
usage: compute-sanitizer --tool memcheck python tools/san_migrate.py [n_envs] [steps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch

import gym_lowcostrobot_b200 as glr
from gym_lowcostrobot_b200.vec import TrajectoryRecorder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
env = glr.make("PushCube-v0", num_envs=n, exec_mode="phased", autoreset=True, max_episode_steps=3)
env.reset(seed=1)
rng = np.random.default_rng(11)
lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])
qpos = np.zeros((n, 13))
qpos[:, :6] = rng.uniform(lo, hi, size=(n, 6))
qpos[:, 1] = rng.uniform(0.9, 1.22, n)  # shoulder forward, elbow down: the arm lies in the floor (100 - 200 constraint rows)
qpos[:, 2] = rng.uniform(1.2, 1.74, n)
qpos[:, 6:9] = [0.0, 0.2, 0.015]
qpos[:, 9] = 1
env.set_state(qpos=qpos, qvel=np.zeros((n, 12)), ctrl=qpos[:, :6], warm=np.zeros((n, 12)))
rec = TrajectoryRecorder(n, env.action_dim, horizon=3, device="cuda:0", pool_episodes=2 * n)
g = torch.Generator(device="cuda").manual_seed(0)
peak = 0
for t in range(steps):
    a = 0.2 * (torch.rand(n, env.action_dim, generator=g, device="cuda") * 2 - 1)
    obs, rew, te, tr, su = env.step_flat(a)
    rec.record(obs, a, te, tr)
    peak = max(peak, int(env.diagnostics()["max_nefc"].max()))
rec.flush()
torch.cuda.synchronize()
print("done: peak rows", peak, "episodes", rec.n_finished, "overflow", int(env.diagnostics()["overflow"].sum()))
