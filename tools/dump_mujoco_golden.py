#!/usr/bin/env python
"""Dump golden vectors from the REAL reference (MuJoCo + gym_lowcostrobot) in the layout of
tests/golden/*.npz.  Not runnable in the build image (mujoco / gymnasium are not installed there);
run it on any machine that has them:

    pip install mujoco gymnasium
    python tools/dump_mujoco_golden.py /path/to/gym-lowcostrobot  out_dir

and drop the files into tests/golden_mujoco/ -- tests/test_golden.py picks them up automatically and
checks the oracle against them (which would turn "parity unpinned" into a pinned oracle).
"""
import os
import sys

import numpy as np


def main(ref_root, out_dir):
    sys.path.insert(0, ref_root)
    import gymnasium as gym
    import gym_lowcostrobot  # noqa: F401
    import mujoco

    ids = {"reach": "ReachCube-v0", "push": "PushCube-v0", "lift": "LiftCube-v0", "pick_place": "PickPlaceCube-v0",
           "stack": "StackTwoCubes-v0", "push_loop": "PushCubeLoop-v0"}
    cases = [("reach", "joint"), ("reach", "ee"), ("push", "joint"), ("lift", "joint"), ("lift", "ee"),
             ("pick_place", "joint"), ("pick_place", "ee"), ("stack", "joint"), ("push_loop", "joint"), ("push_loop", "ee")]
    n_env, n_step = 4, 12
    os.makedirs(out_dir, exist_ok=True)
    for task, mode in cases:
        rng = np.random.default_rng(2024)
        envs = [gym.make(ids[task], observation_mode="state", action_mode=mode).unwrapped for _ in range(n_env)]
        keys = None
        obs0 = []
        for i, e in enumerate(envs):
            o, _ = e.reset(seed=100 + i)
            keys = list(o.keys())
            obs0.append(np.concatenate([o[k] for k in keys]))
        na = envs[0].action_space.shape[0]
        actions = rng.uniform(-1, 1, size=(n_step, n_env, na)).astype(np.float32)
        rec = {k: [] for k in ("obs", "reward", "flags", "qpos", "qvel", "ncon", "nefc")}
        for t in range(n_step):
            row = {k: [] for k in rec}
            for i, e in enumerate(envs):
                o, r, te, tr, info = e.step(actions[t, i])
                row["obs"].append(np.concatenate([o[k] for k in keys]))
                row["reward"].append(r)
                row["flags"].append((te, tr, info.get("is_success", False)))
                row["qpos"].append(e.data.qpos.copy())
                row["qvel"].append(e.data.qvel.copy())
                row["ncon"].append(e.data.ncon)
                row["nefc"].append(e.data.nefc)
            for k in rec:
                rec[k].append(np.array(row[k]))
        np.savez_compressed(os.path.join(out_dir, f"{task}_{mode}.npz"), seed0=100, actions=actions, obs0=np.stack(obs0),
                            obs_keys=np.array(keys), mujoco_version=mujoco.__version__, **{k: np.stack(v) for k, v in rec.items()})
        print(task, mode, "ok")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
