#!/usr/bin/env python
"""GPU diagnostic: PushCubeLoop float64 CUDA vs oracle, substep by substep from synchronised states.
Prints, for every substep whose result differs by more than 1e-9, the env, the contact lists of both sides and the
state difference.  usage: python tools/debug_loop.py [task] [n_env] [n_step]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gym_lowcostrobot_b200 as glr  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

task = sys.argv[1] if len(sys.argv) > 1 else "push_loop"
n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n_step = int(sys.argv[3]) if len(sys.argv) > 3 else 6
IDS = {"reach": "ReachCube-v0", "push": "PushCube-v0", "push_loop": "PushCubeLoop-v0", "stack": "StackTwoCubes-v0"}
env = glr.make(IDS[task], num_envs=n_env, precision="float64")
oracles = [Oracle(task) for _ in range(n_env)]
env.reset(seed=0)
for i, o in enumerate(oracles):
    o.reset(seed=i)
rng = np.random.default_rng(1234)
nq, nv = env.nq, env.nv
reported = 0
for t in range(n_step):
    a = rng.uniform(-1, 1, size=(n_env, env.action_dim)).astype(np.float32)
    # the oracle drives; ctrl as the env would set it (joint mode)
    lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
    hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])
    for i, o in enumerate(oracles):
        st = o.get_state()
        ctrl = np.clip(a[i, :5].astype(np.float64) + st["qpos"][:5], lo[:5], hi[:5])
        o.set_state(ctrl=np.r_[ctrl, 0.0])
    for k in range(20):
        sts = [o.get_state() for o in oracles]
        env.set_state(qpos=np.stack([s["qpos"] for s in sts]), qvel=np.stack([s["qvel"] for s in sts]),
                      ctrl=np.stack([s["ctrl"] for s in sts]), warm=np.stack([s["warm"] for s in sts]))
        con, ncon = env.debug_contacts()
        con, ncon = con.cpu().numpy(), ncon.cpu().numpy()
        env.substeps(1)
        g = {k2: v.cpu().numpy() for k2, v in env.get_state().items()}
        for i, o in enumerate(oracles):
            o.set_state(qpos=sts[i]["qpos"], qvel=sts[i]["qvel"], ctrl=sts[i]["ctrl"], warm=sts[i]["warm"])  # clears the cache like the CUDA side
            o.forward()
            oc = o.get("contacts").reshape(-1, 27)
            o.set_state(qpos=sts[i]["qpos"], qvel=sts[i]["qvel"], ctrl=sts[i]["ctrl"], warm=sts[i]["warm"])
            o.substep(1)
            r = o.get_state()
            dq, dv = np.abs(g["qpos"][i] - r["qpos"]).max(), np.abs(g["qvel"][i] - r["qvel"]).max()
            if (dq > 1e-9 or dv > 1e-6) and reported < 12:
                reported += 1
                print(f"step {t} substep {k} env {i}: dqpos {dq:.3e} dqvel {dv:.3e}; ncon cuda {ncon[i]} oracle {len(oc)}; oracle diag {o.diag()}")
                for c in oc:
                    print("   oracle g1 %2d g2 %2d dim %d dist %+.6e pos %s n %s" % (c[14], c[15], c[13], c[12], np.round(c[0:3], 6), np.round(c[3:6], 6)))
                for c in con[i, :ncon[i]]:
                    print("   cuda   b1 %2d b2 %2d dim %d dist %+.6e pos %s n %s" % (c[7], c[8], c[9], c[6], np.round(c[0:3], 6), np.round(c[3:6], 6)))
                print("   dqvel", np.round(g["qvel"][i] - r["qvel"], 9))
print("reported", reported)
