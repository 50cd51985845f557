import sys, time; sys.path.insert(0, '/root/repo')
import torch, gym_lowcostrobot_b200 as glr
for n in (256, 4096):
    env = glr.make("PushCube-v0", num_envs=n, observation_mode="both", autoreset=True)
    env.reset(seed=0)
    a = torch.rand(n, env.action_dim, device="cuda") * 2 - 1
    for _ in range(3): env.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): env._renderer.render()
    e1.record(); torch.cuda.synchronize()
    print(f"n={n}: render 2 x 240x320 per env: {e0.elapsed_time(e1)/5:.3f} ms per batch, {n*2/(e0.elapsed_time(e1)/5)*1e3:.0f} images/s")
    env.close()
