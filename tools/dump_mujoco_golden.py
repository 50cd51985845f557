#!/usr/bin/env python
"""Dump golden vectors from the REAL reference (MuJoCo + gym_lowcostrobot) in the layout of
tests/golden/*.npz.  Not runnable in the build image (mujoco / gymnasium are not installed there);
run it on any machine that has them:

    pip install mujoco gymnasium
    python tools/dump_mujoco_golden.py /path/to/gym-lowcostrobot  out_dir

and drop the files into tests/golden_mujoco/ -- tests/test_golden_mujoco.py picks them up automatically and checks the
oracle (CPU) and the CUDA path (GPU) against them, which turns "parity unpinned" into a pinned oracle without code changes.
Besides the rollouts of tests/golden/*.npz the files carry what a mismatch is bisected with: the compiled model constants
(opt.impratio / timestep, body_invweight0, dof_invweight0), and for env 0 the state before every env.step together with the
result of ONE mujoco.mj_step from it (qpos, qvel, ncon, nefc and the contact list: pos, normal, dist, geom ids, dim).
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
        sub = {k: [] for k in ("sub_qpos0", "sub_qvel0", "sub_ctrl", "sub_warm", "sub_qpos1", "sub_qvel1", "sub_ncon", "sub_nefc", "sub_contacts")}
        m0 = envs[0].model
        for t in range(n_step):
            # one mj_step from env 0's current state, on a copy (the ctrl of the previous env.step is still set)
            d0 = envs[0].data
            d = mujoco.MjData(m0)
            d.qpos[:], d.qvel[:], d.ctrl[:], d.qacc_warmstart[:], d.time = d0.qpos, d0.qvel, d0.ctrl, d0.qacc_warmstart, d0.time
            for k, v in (("sub_qpos0", d.qpos), ("sub_qvel0", d.qvel), ("sub_ctrl", d.ctrl), ("sub_warm", d.qacc_warmstart)):
                sub[k].append(np.array(v))
            mujoco.mj_step(m0, d)
            con = np.zeros((128, 10))
            for c in range(min(d.ncon, 128)):
                ct = d.contact[c]
                con[c] = np.r_[ct.pos, ct.frame[:3], ct.dist, ct.geom1, ct.geom2, ct.dim]
            for k, v in (("sub_qpos1", d.qpos), ("sub_qvel1", d.qvel), ("sub_ncon", d.ncon), ("sub_nefc", d.nefc), ("sub_contacts", con)):
                sub[k].append(np.array(v))
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
                            obs_keys=np.array(keys), mujoco_version=mujoco.__version__, opt_impratio=m0.opt.impratio, opt_timestep=m0.opt.timestep,
                            body_invweight0=np.array(m0.body_invweight0), dof_invweight0=np.array(m0.dof_invweight0), body_mass=np.array(m0.body_mass),
                            **{k: np.stack(v) for k, v in rec.items()}, **{k: np.stack(v) for k, v in sub.items()})
        print(task, mode, "ok")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
