"""Step-by-step GPU bring-up diagnostic: prints after every stage so that a hang is localised."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
t0 = time.time()


def log(*a):
    print(f"[{time.time() - t0:7.2f}s]", *a, flush=True)


import numpy as np  # noqa: E402
import torch  # noqa: E402

log("torch", torch.__version__, torch.cuda.get_device_name(0))
import gym_lowcostrobot_b200 as glr  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

np.set_printoptions(precision=6, suppress=True, linewidth=200)
task = sys.argv[1] if len(sys.argv) > 1 else "reach"
prec = sys.argv[2] if len(sys.argv) > 2 else "float64"
mode = sys.argv[3] if len(sys.argv) > 3 else "joint"
IDS = {"reach": "ReachCube-v0", "push": "PushCube-v0", "lift": "LiftCube-v0", "pick_place": "PickPlaceCube-v0", "stack": "StackTwoCubes-v0"}
n = 4
env = glr.make(IDS[task], num_envs=n, precision=prec, action_mode=mode)
log("created")
obs, _ = env.reset(seed=0)
torch.cuda.synchronize()
log("reset", {k: v[0].cpu().numpy() for k, v in obs.items()})
orc = [Oracle(task, action_mode=mode) for _ in range(n)]
for i, o in enumerate(orc):
    o.reset(seed=i)
st = env.get_state()
torch.cuda.synchronize()
log("state qpos0", st["qpos"][0].cpu().numpy())
log("oracle qpos0", orc[0].get_state()["qpos"])
log("warm gpu", st["warm"][0].cpu().numpy())
log("warm orc", orc[0].get_state()["warm"])
log("diag", {k: v.cpu().numpy() for k, v in env.diagnostics().items()}, orc[0].diag())
env.substeps(1)
torch.cuda.synchronize()
for o in orc:
    o.substep(1)
st = env.get_state()
log("after 1 substep: max|dqpos|", np.abs(st["qpos"].cpu().numpy() - np.stack([o.get_state()["qpos"] for o in orc])).max(),
    "max|dqvel|", np.abs(st["qvel"].cpu().numpy() - np.stack([o.get_state()["qvel"] for o in orc])).max())
rng = np.random.default_rng(0)
for t in range(5):
    a = rng.uniform(-1, 1, size=(n, env.action_dim)).astype(np.float32)
    ob, r, te, tr, info = env.step(torch.from_numpy(a).cuda())
    torch.cuda.synchronize()
    ref = [o.step(a[i]) for i, o in enumerate(orc)]
    st = env.get_state()
    dq = np.abs(st["qpos"].cpu().numpy() - np.stack([o.get_state()["qpos"] for o in orc])).max()
    dv = np.abs(st["qvel"].cpu().numpy() - np.stack([o.get_state()["qvel"] for o in orc])).max()
    log(f"step {t}: max|dqpos| {dq:.3e} max|dqvel| {dv:.3e} reward", r.cpu().numpy(), [x[1] for x in ref],
        "diag", {k: v.cpu().numpy().tolist() for k, v in env.diagnostics().items()}, orc[0].diag())
log("done")
