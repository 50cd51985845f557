"""Env constructor kwargs -> ``LcrEnvCfg`` (mirrors the reference ``__init__`` signatures).

Reference: ``reach_cube_env.py:77-139``, ``push_cube_env.py:79-148``, ``lift_cube_env.py:77-144``,
``pick_place_cube_env.py:79-153``, ``stack_two_cubes_env.py:78-144`` and the TimeLimit of
``gym_lowcostrobot/__init__.py:9-37`` (``max_episode_steps=50``).
"""
from __future__ import annotations

from .model import COLLIDE_ALL, LcrEnvCfg

ENV_IDS = {
    "ReachCube-v0": "reach",
    "PushCube-v0": "push",
    "LiftCube-v0": "lift",
    "PickPlaceCube-v0": "pick_place",
    "StackTwoCubes-v0": "stack",
    "PushCubeLoop-v0": "push_loop",  # push_cube_loop_env.py, registered at gym_lowcostrobot/__init__.py:39-43
}
# reference defaults of block_gripper per task
BLOCK_GRIPPER_DEFAULT = {"reach": True, "push": True, "lift": False, "pick_place": False, "stack": False, "push_loop": True}
MAX_EPISODE_STEPS = 50


def make_cfg(task, action_mode="joint", reward_type="sparse", block_gripper=None, distance_threshold=0.05,
             height_threshold=0.1, cube_xy_range=0.3, target_xy_range=0.3, goal_z_range=0.1, n_substeps=20,
             max_episode_steps=MAX_EPISODE_STEPS, autoreset=False, collision_mask=COLLIDE_ALL, exec_mode=0):
    if action_mode not in ("joint", "ee"):
        raise ValueError("Invalid action mode, must be 'ee' or 'joint'")  # reach_cube_env.py:270
    if reward_type not in ("sparse", "dense"):
        raise ValueError("reward_type must be 'sparse' or 'dense'")
    if block_gripper is None:
        block_gripper = BLOCK_GRIPPER_DEFAULT[task]
    cfg = LcrEnvCfg()
    cfg.action_mode = 1 if action_mode == "ee" else 0
    cfg.block_gripper = int(bool(block_gripper))
    cfg.reward_type = 1 if reward_type == "dense" else 0
    cfg.n_substeps = int(n_substeps)
    cfg.max_episode_steps = int(max_episode_steps) if max_episode_steps else 0
    cfg.autoreset = int(bool(autoreset))
    cfg.collision_mask = int(collision_mask)
    cfg.exec_mode = int(exec_mode)
    cfg.distance_threshold = float(distance_threshold)
    cfg.height_threshold = float(height_threshold)
    # sampling boxes (reach_cube_env.py:134-139): xy range centred, then y shifted
    lo = [-cube_xy_range / 2, -cube_xy_range / 2 + 0.165, 0.0]
    hi = [cube_xy_range / 2, cube_xy_range / 2 + 0.10, 0.0]
    tlo = [-target_xy_range / 2, -target_xy_range / 2 + 0.165, 0.0]
    thi = [target_xy_range / 2, target_xy_range / 2 + 0.10, goal_z_range if task == "pick_place" else 0.0]
    for k in range(3):
        cfg.cube_low[k], cfg.cube_high[k] = lo[k], hi[k]
        cfg.target_low[k], cfg.target_high[k] = tlo[k], thi[k]
    return cfg


def action_dim(cfg):
    return (3 if cfg.action_mode else 5) + (0 if cfg.block_gripper else 1)


def obs_dim(task):
    return 15 if task in ("reach", "lift", "push_loop") else 18
