#!/usr/bin/env python
"""GPU diagnostic: (1) the one-substep map of tests/test_gpu_parity.py for push_loop, details of the envs that differ;
(2) free-running 120-substep rollouts (persistent caches), growth of the CUDA-vs-oracle difference per env."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import gym_lowcostrobot_b200 as glr  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

spec = importlib.util.spec_from_file_location("tp", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
tp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tp)

task = "push_loop"
n = 192
con, ncon, ref_con = tp._contact_lists(task, "float64", 127, n=n)
env = glr.make("PushCubeLoop-v0", num_envs=n, precision="float64")
rng = np.random.default_rng(7)
qpos, qvel, ctrl = tp.random_states(task, n, rng, env.nq, env.nv)
env.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, warm=np.zeros((n, env.nv)))
env.substeps(1)
st = {k: v.cpu().numpy() for k, v in env.get_state().items()}
dg = {k: v.cpu().numpy() for k, v in env.diagnostics().items()}
worst = []
for i in range(n):
    same = len(ref_con[i]) == ncon[i] and tp._contact_err(con[i, :ncon[i]], ref_con[i]) <= 1e-9
    o = Oracle(task)
    o.set_state(qpos=qpos[i], qvel=qvel[i], ctrl=ctrl[i], warm=np.zeros(env.nv))
    o.substep(1)
    ref = o.get_state()
    dq, dv = np.abs(st["qpos"][i] - ref["qpos"]).max(), np.abs(st["qvel"][i] - ref["qvel"]).max()
    worst.append((dv, dq, i, same, o.diag(), {k: int(v[i]) for k, v in dg.items()}, np.abs(ref["qvel"]).max()))
worst.sort(key=lambda x: -x[0])
print("one-substep map, 8 worst envs by |dqvel|:")
for w in worst[:8]:
    print("  dqvel %.3e dqpos %.3e env %d same_contacts %s\n     oracle %s\n     cuda   %s  max|qvel| %.3e" % w)
i = worst[0][2]
print("contacts of the worst env (oracle):")
for c in ref_con[i]:
    print("   g1 %2d g2 %2d dim %d dist %+.6e" % (c[14], c[15], c[13], c[12]))
env.close()

# (2) free-running rollouts exactly as tests/test_gpu_parity.py::rollout_pair, difference after every env step
n_env, n_step = 16, 6
env = glr.make("PushCubeLoop-v0", num_envs=n_env, precision="float64")
oracles = [Oracle(task) for _ in range(n_env)]
env.reset(seed=0)
for k, o in enumerate(oracles):
    o.reset(seed=k)
rng = np.random.default_rng(1234)
for t in range(n_step):
    a = rng.uniform(-1, 1, size=(n_env, env.action_dim)).astype(np.float32)
    env.step(torch.from_numpy(a).cuda())
    g = {k: v.cpu().numpy() for k, v in env.get_state().items()}
    d = {k: v.cpu().numpy() for k, v in env.diagnostics().items()}
    line = []
    for k, o in enumerate(oracles):
        o.step(a[k])
        r = o.get_state()
        od = o.diag()
        line.append("%d:%.1e/%d,%d,%d|%d,%d,%d" % (k, np.abs(g["qpos"][k] - r["qpos"]).max(), od["ncon"], od["max_nefc"], od["niter"],
                                                 d["ncon"][k], d["max_nefc"][k], d["niter"][k]))
    print("step", t, " ".join(line))
