"""Debug: one-substep map from random contact-rich states, CUDA f64 vs oracle; prints per-env mismatch details."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import gym_lowcostrobot_b200 as glr
from oracle.oracle import Oracle
from test_gpu_parity import random_states, IDS
np.set_printoptions(precision=6, suppress=True, linewidth=220)
task = sys.argv[1] if len(sys.argv) > 1 else "push"
prec = sys.argv[2] if len(sys.argv) > 2 else "float64"
mask = int(sys.argv[3]) if len(sys.argv) > 3 else 31
n = 192
env = glr.make(IDS[task], num_envs=n, precision=prec, collision_mask=mask)
rng = np.random.default_rng(7)
qpos, qvel, ctrl = random_states(task, n, rng, env.nq, env.nv)
env.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, warm=np.zeros((n, env.nv)))
env.substeps(1)
st = {k: v.cpu().numpy() for k, v in env.get_state().items()}
diag = {k: v.cpu().numpy() for k, v in env.diagnostics().items()}
nbad = 0
for i in range(n):
    o = Oracle(task, collision_mask=mask)
    o.set_state(qpos=qpos[i], qvel=qvel[i], ctrl=ctrl[i], warm=np.zeros(env.nv))
    o.substep(1)
    ref = o.get_state(); d = o.diag()
    eq = np.abs(st["qpos"][i] - ref["qpos"]).max(); ev = np.abs(st["qvel"][i] - ref["qvel"]).max()
    same = d["ncon"] == diag["ncon"][i] and d["nefc"] == diag["nefc"][i]
    if not same or eq > 1e-7 or ev > 1e-4:
        nbad += 1
        if nbad <= 12:
            con = o.get("contacts").reshape(-1, 27)
            print(f"env {i}: oracle ncon {d['ncon']} nefc {d['nefc']} niter {d['niter']} | gpu ncon {diag['ncon'][i]} nefc {diag['nefc'][i]} niter {diag['niter'][i]} | dq {eq:.2e} dv {ev:.2e}")
            print("   oracle contacts (g1,g2,dist):", [(int(c[14]), int(c[15]), float(f"{c[12]:.3e}")) for c in con])
print("mismatching envs:", nbad, "of", n)
