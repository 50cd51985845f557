"""Debug: compare contact lists (geometry) of mj_forward, CUDA f64 vs oracle, from random states."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import gym_lowcostrobot_b200 as glr
from oracle.oracle import Oracle
from test_gpu_parity import random_states, IDS
np.set_printoptions(precision=6, suppress=True, linewidth=220)
task = sys.argv[1]; prec = sys.argv[2]; mask = int(sys.argv[3]); n = 192
env = glr.make(IDS[task], num_envs=n, precision=prec, collision_mask=mask)
rng = np.random.default_rng(7)
qpos, qvel, ctrl = random_states(task, n, rng, env.nq, env.nv)
env.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, warm=np.zeros((n, env.nv)))
con, ncon = env.debug_contacts()
con, ncon = con.cpu().numpy(), ncon.cpu().numpy()
nb = 0; worst = 0
for i in range(n):
    o = Oracle(task, collision_mask=mask)
    o.set_state(qpos=qpos[i], qvel=qvel[i], ctrl=ctrl[i], warm=np.zeros(env.nv))
    o.forward()
    oc = o.get("contacts").reshape(-1, 27)
    g = con[i, :ncon[i]]
    ok = len(oc) == ncon[i]
    err = 0
    if ok and len(oc):
        err = max(np.abs(oc[:, 0:3] - g[:, 0:3]).max(), np.abs(oc[:, 3:6] - g[:, 3:6]).max(), np.abs(oc[:, 12] - g[:, 6]).max())
        worst = max(worst, err)
    if not ok or err > 1e-9:
        nb += 1
        if nb <= 6:
            print(f"env {i}: oracle ncon {len(oc)} gpu ncon {ncon[i]} err {err:.2e}")
            print("  oracle:", np.c_[oc[:, 14:16], oc[:, 12], oc[:, 0:6]])
            print("  gpu   :", np.c_[g[:, 7:9], g[:, 6], g[:, 0:6]])
print("mismatch", nb, "of", n, "worst err among matching counts", worst)
