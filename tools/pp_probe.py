import sys, time; sys.path.insert(0, '/root/repo')
import torch, bench
off = int(sys.argv[1]) if len(sys.argv) > 1 else 0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = bench.GpuRun("pick_place", 8192, "ee", "auto", off, 2 if off else 1, 0, 10, 3) if False else None
import gym_lowcostrobot_b200 as glr
n = 8192
env = glr.make("PickPlaceCube-v0", num_envs=n, action_mode="ee", autoreset=True, env_offset=off * n)
env.reset(seed=0)
ints = torch.zeros(n, 2, dtype=torch.int32, device="cuda"); ints[:, 0] = (torch.arange(n, device="cuda") + off * n) % 50
env.set_state(ints=ints)
gen = torch.Generator(device="cuda").manual_seed(1234 + off)
rec = torch.empty(n, env.obs_dim + 4, device="cuda")
out = []
for t in range(150):
    a = torch.rand(n, env.action_dim, generator=gen, device="cuda") * 2 - 1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    env.step_packed(a, out=rec)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    if t >= 100:
        d = env.diagnostics()
        out.append((round(dt, 1), int((d["max_nefc"] >= 97).sum()), int(d["max_nefc"].max())))
print("offset", off, out)
