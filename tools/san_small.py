"""Tiny runs for compute-sanitizer: usage san_small.py <env id> <exec mode> [n_envs] [steps]
(all entry points that write caller buffers: reset with a mask, step, pack, get/set state, ik, debug contacts)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import gym_lowcostrobot_b200 as glr
env_id, mode = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 24
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
env = glr.make(env_id, num_envs=n, autoreset=True, exec_mode=mode, max_episode_steps=2)
env.reset(seed=1)
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(steps):
    env.step_packed(torch.rand(n, env.action_dim, generator=g, device="cuda") * 2 - 1)
mask = torch.zeros(n, dtype=torch.bool, device="cuda")
mask[::3] = True
env.reset(mask=mask)
st = env.get_state()
env.set_state(**st)
env.inverse_kinematics(torch.rand(n, 3, generator=g, device="cuda") * 0.2)
env.debug_contacts()
env.diagnostics()
torch.cuda.synchronize()
print("done", env_id, mode)
