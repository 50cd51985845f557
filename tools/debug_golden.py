#!/usr/bin/env python
"""Debug helper: CUDA float64 vs oracle on a golden case, step by step (diag + contact lists at the first mismatch)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import gym_lowcostrobot_b200 as glr
from oracle.oracle import Oracle
case = sys.argv[1] if len(sys.argv) > 1 else "reach_joint"
mode_exec = sys.argv[2] if len(sys.argv) > 2 else "auto"
task, mode = case.rsplit("_", 1)
IDS = {"reach": "ReachCube-v0", "push": "PushCube-v0", "lift": "LiftCube-v0", "pick_place": "PickPlaceCube-v0", "stack": "StackTwoCubes-v0"}
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", case + ".npz"))
n_step, n_env, _ = z["actions"].shape
env = glr.make(IDS[task], num_envs=n_env, action_mode=mode, precision="float64", exec_mode=mode_exec)
env.reset(seed=int(z["seed0"]))
orc = [Oracle(task, action_mode=mode) for _ in range(n_env)]
for i, o in enumerate(orc): o.reset(seed=int(z["seed0"]) + i)
def contacts(i):
    c, nc = env.debug_contacts()
    c = c[i, :int(nc[i])].cpu().numpy()
    orc[i].forward()
    oc = orc[i].get("contacts").reshape(-1, 27)
    return c, oc
for t in range(n_step):
    env.step(torch.from_numpy(z["actions"][t]).cuda())
    q = env.get_state()["qpos"].cpu().numpy()
    dg = {k: v.cpu().numpy() for k, v in env.diagnostics().items()}
    bad = []
    for i, o in enumerate(orc):
        o.step(z["actions"][t, i])
        err = np.abs(o.get_state()["qpos"] - q[i]).max()
        od = o.diag()
        print(f"t={t} env={i} |dq|={err:.2e}  cuda ncon/nefc/niter/max={dg['ncon'][i]}/{dg['nefc'][i]}/{dg['niter'][i]}/{dg['max_nefc'][i]}  oracle {od['ncon']}/{od['nefc']}/{od['niter']}/{od['max_nefc']}")
        if err > 1e-7: bad.append(i)
    if bad:
        for i in bad:
            c, oc = contacts(i)
            print("env", i, "CUDA contacts (pos, normal, dist, b1, b2):")
            for r in c: print("   ", np.round(r[:3], 5), np.round(r[3:6], 4), f"{r[6]:.3e}", int(r[7]), int(r[8]))
            print("env", i, "oracle contacts (pos, normal, dist, g1, g2):")
            for r in oc: print("   ", np.round(r[:3], 5), np.round(r[3:6], 4), f"{r[12]:.3e}", int(r[14]), int(r[15]))
        break
