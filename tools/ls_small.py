"""Tiny lockstep run (for compute-sanitizer): 24 Stack envs, 3 steps."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import gym_lowcostrobot_b200 as glr
env = glr.make("StackTwoCubes-v0", num_envs=24, autoreset=True, exec_mode="lockstep")
env.reset(seed=1)
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(3):
    env.step(torch.rand(24, env.action_dim, generator=g, device="cuda") * 2 - 1)
torch.cuda.synchronize()
print("done")
