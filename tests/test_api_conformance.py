"""API conformance in the spirit of gymnasium's check_env (the reference's only test, tests/test_env.py:8-13), batched:
declared spaces, observation keys / shapes / dtypes / bounds, the 5-tuple of step, info keys per env, error behaviour
(reference reach_cube_env.py:95-115,231-232,281-295,313-333; lift_cube_env.py:337; push_cube_loop_env.py:320,329)."""
import numpy as np
import pytest
import torch

import gym_lowcostrobot_b200 as glr

pytestmark = pytest.mark.gpu

CASES = [  # env id, observation keys in the reference's order, default block_gripper
    ("ReachCube-v0", ("arm_qpos", "arm_qvel", "cube_pos"), True),
    ("PushCube-v0", ("arm_qpos", "arm_qvel", "target_pos", "cube_pos"), True),
    ("LiftCube-v0", ("arm_qpos", "arm_qvel", "cube_pos"), False),
    ("PickPlaceCube-v0", ("arm_qpos", "arm_qvel", "target_pos", "cube_pos"), False),
    ("StackTwoCubes-v0", ("arm_qpos", "arm_qvel", "cube_red_pos", "cube_blue_pos"), False),
    ("PushCubeLoop-v0", ("arm_qpos", "arm_qvel", "cube_pos"), True),
]


@pytest.mark.parametrize("mode", ["joint", "ee"])
@pytest.mark.parametrize("env_id,keys,blocked", CASES, ids=[c[0] for c in CASES])
def test_env_api(env_id, keys, blocked, mode):
    n = 3
    env = glr.make(env_id, num_envs=n, action_mode=mode)
    na = {"joint": 5, "ee": 3}[mode] + (0 if blocked else 1)
    assert env.single_action_space.shape == (na,) and env.single_action_space.dtype == np.float32
    assert float(env.single_action_space.low[0]) == -1.0 and float(env.single_action_space.high[0]) == 1.0
    assert tuple(env.single_observation_space) == keys
    assert env.action_space.shape == (n, na)

    def check_obs(obs):
        assert tuple(obs) == keys
        for k in keys:
            w = 6 if k.startswith("arm") else 3
            assert obs[k].shape == (n, w) and obs[k].dtype == torch.float32 and obs[k].is_cuda
            assert env.observation_space[k].contains(obs[k].cpu().numpy()), k

    obs, info = env.reset(seed=0)
    check_obs(obs)
    assert info == ({"timestamp": 0.0} if env_id == "PushCubeLoop-v0" else {})
    assert torch.all(obs["arm_qpos"] == 0)  # reset puts the arm at q = 0 (reach_cube_env.py:304-305)
    g = torch.Generator(device="cuda").manual_seed(1)
    first = None
    for t in range(50):
        a = (torch.rand(n, na, generator=g, device="cuda") * 2 - 1) * 0.15
        out = env.step(a)
        assert len(out) == 5
        obs2, reward, terminated, truncated, info = out
        if first is None:
            first = {k: v.clone() for k, v in obs2.items()}
            kept = obs2
        assert reward.shape == (n,) and reward.dtype == torch.float32
        assert terminated.shape == (n,) and terminated.dtype == torch.bool and truncated.dtype == torch.bool
        if env_id == "LiftCube-v0":
            assert info == {} and not terminated.any()
        elif env_id == "PushCubeLoop-v0":
            assert set(info) == {"timestamp", "success"} and not terminated.any()
            assert info["success"].shape == (n,) and torch.allclose(info["timestamp"], torch.full((n,), 0.04 * (t + 1), dtype=torch.float64, device="cuda"), rtol=0, atol=2e-4)
        else:
            assert set(info) == {"is_success"} and info["is_success"].dtype == torch.bool
            assert torch.equal(info["is_success"], terminated)
        assert torch.equal(truncated, torch.full((n,), t == 49, device="cuda"))  # TimeLimit(50) of the registration
    check_obs(obs2)
    for k in keys:  # returned observations are fresh copies, like the reference's astype(np.float32)
        assert torch.equal(kept[k], first[k])
    with pytest.raises(ValueError, match="Action dimension mismatch"):
        env.step(torch.zeros(n, na + 1, device="cuda"))
    with pytest.raises(ValueError, match="Action dimension mismatch"):
        env.step(np.zeros((na,), np.float32))
    env.close()


def test_failure_reporting_in_info():
    """per-env blow-up flag and counters in info, on request (the default info carries exactly the reference's keys)"""
    for env_id in ("PushCube-v0", "PushCubeLoop-v0"):
        env = glr.make(env_id, num_envs=16, report_failures=True)
        env.reset(seed=0)
        obs, reward, terminated, truncated, info = env.step(torch.zeros(16, env.action_dim, device="cuda"))
        assert {"nan_reset", "nan_resets", "contact_overflow"} <= set(info)
        assert info["nan_reset"].dtype == torch.bool and not info["nan_reset"].any() and int(info["contact_overflow"].sum()) == 0
        # a poisoned state is detected, reset like mj_checkPos does, flagged once and counted
        q = env.get_state()["qpos"]
        q[3, 0] = float("nan")
        env.set_state(qpos=q)
        _, _, _, _, info = env.step(torch.zeros(16, env.action_dim, device="cuda"))
        assert info["nan_reset"].tolist() == [i == 3 for i in range(16)] and int(info["nan_resets"][3]) >= 1
        _, _, _, _, info = env.step(torch.zeros(16, env.action_dim, device="cuda"))
        assert not info["nan_reset"].any() and int(info["nan_resets"][3]) >= 1
        env.close()


def test_constructor_errors():
    with pytest.raises(ValueError):
        glr.make("ReachCube-v0", num_envs=2, observation_mode="depth")
    with pytest.raises(NotImplementedError):
        glr.make("PushCubeLoop-v0", num_envs=2, render_mode="human")
    with pytest.raises(ValueError, match="Invalid action mode"):
        glr.make("LiftCube-v0", num_envs=2, action_mode="torque")
    with pytest.raises(TypeError):
        glr.make("PushCubeLoop-v0", num_envs=2, reward_type="dense")  # not in the reference signature
    with pytest.raises(KeyError):
        glr.make("NoSuchTask-v0")
