"""compute-sanitizer target: a few steps of one execution mode on a small batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gym_lowcostrobot_b200 as glr

mode = sys.argv[1] if len(sys.argv) > 1 else "flow"
task = sys.argv[2] if len(sys.argv) > 2 else "ReachCube-v0"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 67
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
env = glr.make(task, num_envs=n, autoreset=True, max_episode_steps=3, exec_mode=mode)
env.reset(seed=3)
gen = torch.Generator(device="cuda").manual_seed(0)
for t in range(steps):
    a = torch.rand(n, env.action_dim, generator=gen, device="cuda") * 2 - 1
    env.step_packed(a)
    torch.cuda.synchronize()
    print("step", t, "ok", (env.flow_status(), env.flow_debug) if mode == "flow" else "", flush=True)
env.close()
