#!/usr/bin/env python
"""Generate tests/golden/*.npz: short seeded rollouts of every task / action mode through the float64
oracle (oracle/lcr_oracle.c).  MuJoCo is not installable in the build image, so these fixtures pin the
ORACLE (regression) and the CUDA path (parity); tools/dump_mujoco_golden.py produces the same layout
from real MuJoCo wherever it is available.

Usage: python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle  # noqa: E402

CASES = [("reach", "joint"), ("reach", "ee"), ("push", "joint"), ("lift", "joint"), ("lift", "ee"),
         ("pick_place", "joint"), ("pick_place", "ee"), ("stack", "joint"), ("push_loop", "joint"), ("push_loop", "ee")]
N_ENV, N_STEP = 4, 12

out_dir = os.path.join(ROOT, "tests", "golden")
os.makedirs(out_dir, exist_ok=True)
for task, mode in CASES:
    rng = np.random.default_rng(2024)
    envs = [Oracle(task, action_mode=mode) for _ in range(N_ENV)]
    obs0 = np.stack([e.reset(seed=100 + i) for i, e in enumerate(envs)])
    na = envs[0].na
    actions = rng.uniform(-1, 1, size=(N_STEP, N_ENV, na)).astype(np.float32)
    obs = np.zeros((N_STEP, N_ENV, envs[0].no), np.float32)
    rew = np.zeros((N_STEP, N_ENV), np.float32)
    flags = np.zeros((N_STEP, N_ENV, 3), np.uint8)
    qpos = np.zeros((N_STEP, N_ENV, envs[0].nq))
    qvel = np.zeros((N_STEP, N_ENV, envs[0].nv))
    for t in range(N_STEP):
        for i, e in enumerate(envs):
            o, r, te, tr, su = e.step(actions[t, i])
            obs[t, i], rew[t, i], flags[t, i] = o, r, (te, tr, su)
            st = e.get_state()
            qpos[t, i], qvel[t, i] = st["qpos"], st["qvel"]
    path = os.path.join(out_dir, f"{task}_{mode}.npz")
    np.savez_compressed(path, seed0=100, actions=actions, obs0=obs0, obs=obs, reward=rew, flags=flags, qpos=qpos, qvel=qvel)
    print(path, os.path.getsize(path))
