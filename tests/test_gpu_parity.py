"""CUDA path (through the C-ABI) vs the CPU float64 oracle on identical seeds and actions.

float64 build of the kernels: implementation parity, tolerance 1e-8 on states after a full env step
(the two sides run the same algorithm with different summation orders and an iterative solver that
stops at 1e-8).  float32 product build: tolerance stated per test.
"""
import numpy as np
import pytest
import torch

import gym_lowcostrobot_b200 as glr
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

IDS = {"reach": "ReachCube-v0", "push": "PushCube-v0", "lift": "LiftCube-v0", "pick_place": "PickPlaceCube-v0",
       "stack": "StackTwoCubes-v0"}


def rollout_pair(task, action_mode, precision, n_env, n_step, seed=0, act_scale=1.0, **kw):
    env = glr.make(IDS[task], num_envs=n_env, action_mode=action_mode, precision=precision, **kw)
    oracles = [Oracle(task, action_mode=action_mode, **kw) for _ in range(n_env)]
    obs, _ = env.reset(seed=seed)
    flat = torch.cat([obs[k] for k in obs], 1).cpu().numpy()
    for i, o in enumerate(oracles):
        oo = o.reset(seed=seed + i)
        np.testing.assert_array_equal(oo[12:], flat[i, 12:])  # sampled positions are bit-exact
    rng = np.random.default_rng(1234)
    out = []
    for t in range(n_step):
        a = (act_scale * rng.uniform(-1, 1, size=(n_env, env.action_dim))).astype(np.float32)
        o_gpu, r_gpu, te, tr, info = env.step(torch.from_numpy(a).cuda())
        st = {k: v.cpu().numpy() for k, v in env.get_state().items()}
        ref = [o.step(a[i]) for i, o in enumerate(oracles)]
        ref_st = [o.get_state() for o in oracles]
        out.append((st, ref_st, r_gpu.cpu().numpy(), np.array([r[1] for r in ref]), te.cpu().numpy(),
                    np.array([r[2] for r in ref]), tr.cpu().numpy(), np.array([r[3] for r in ref])))
    diag = {k: v.cpu().numpy() for k, v in env.diagnostics().items()}
    env.close()
    return out, diag, oracles


@pytest.mark.parametrize("task", ["reach", "push", "lift", "pick_place", "stack"])
def test_f64_joint_rollout_matches_oracle(task):
    out, diag, oracles = rollout_pair(task, "joint", "float64", n_env=16, n_step=6)
    for st, ref_st, r, r_ref, te, te_ref, tr, tr_ref in out:
        for key in ("qpos", "qvel", "ctrl"):
            ref = np.stack([s[key] for s in ref_st])
            np.testing.assert_allclose(st[key], ref, rtol=0, atol=1e-7, err_msg=f"{task} {key}")
        np.testing.assert_allclose(r, r_ref, atol=1e-6)
        np.testing.assert_array_equal(te, te_ref)
        np.testing.assert_array_equal(tr, tr_ref)


@pytest.mark.parametrize("task", ["reach", "pick_place"])
def test_f64_ee_rollout_matches_oracle(task):
    out, diag, oracles = rollout_pair(task, "ee", "float64", n_env=8, n_step=4)
    for st, ref_st, r, r_ref, te, te_ref, tr, tr_ref in out:
        for key in ("qpos", "qvel", "ctrl"):
            ref = np.stack([s[key] for s in ref_st])
            np.testing.assert_allclose(st[key], ref, rtol=0, atol=1e-7, err_msg=f"{task} {key}")


def test_f32_reach_rollout_close_to_oracle():
    # float32 product path: joint angles within 2e-3 rad and cube position within 1e-3 m of the float64
    # oracle after 5 env steps (100 substeps) of random actions.
    out, diag, oracles = rollout_pair("reach", "joint", "float32", n_env=64, n_step=5)
    st, ref_st = out[-1][0], out[-1][1]
    ref_q = np.stack([s["qpos"] for s in ref_st])
    err_arm = np.abs(st["qpos"][:, :6] - ref_q[:, :6]).max(1)
    err_cube = np.abs(st["qpos"][:, 6:9] - ref_q[:, 6:9]).max(1)
    assert np.median(err_arm) < 2e-3 and np.median(err_cube) < 1e-3, (err_arm, err_cube)
    assert (err_arm < 2e-3).mean() > 0.9
