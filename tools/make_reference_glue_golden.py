#!/usr/bin/env python
"""Generate tests/golden/ref_glue/<task>_<mode>.npz by EXECUTING THE REFERENCE'S OWN ENV CLASSES.

The arithmetic of `mj_step` lives in the third-party `mujoco` package, which cannot be installed in the build
image -- but everything AROUND it is the reference's own Python: action clipping and joint / end-effector action
maps, the damped-least-squares IK loop (`np.linalg.inv`, `np.linalg.pinv`), `get_observation`, `reset` sampling from
gymnasium's seeded generator, success / reward (`compute_reward`, PushCubeLoop's `get_reward` / `get_cube_overlap`
with its goal switching).  This script imports the UNMODIFIED modules from /root/reference/gym_lowcostrobot/envs with
stand-ins for the two absent third-party packages:

  * `gymnasium`: `Env.reset(seed)` seeds `np_random = np.random.default_rng(seed)` (what gymnasium's
    `seeding.np_random` does), `spaces.Box` / `spaces.Dict` hold low / high / shape;
  * `mujoco`: `MjModel.from_xml_path` returns the numbers the MJCF holds (parsed by gym_lowcostrobot_b200.mjcf),
    `mj_forward` is an independent numpy forward-kinematics pass (mjcf.arm_kinematics; no collision, no dynamics),
    `mj_jacSite` the textbook site Jacobian (axis x (p_site - anchor)), `mj_step` is never reached because the envs
    are built with `n_substeps=0`.

So the fixtures pin the GLUE of the hot path to the reference's own code; the
physics inside `mj_step` (row a7) stays pinned by the analytic tests only.  tests/test_reference_glue.py replays the
fixtures through the oracle (CPU) and through the CUDA path (GPU) with `n_substeps=0`.

Usage: python tools/make_reference_glue_golden.py [REFERENCE_ROOT]      (default /root/reference)
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from gym_lowcostrobot_b200 import mjcf  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ASSETS = os.path.join(REF, "gym_lowcostrobot", "assets", "low_cost_robot_6dof")
XML_TASK = {v: k for k, v in mjcf.TASK_XML.items()}


# ------------------------------------------------------------------ stand-in for gymnasium
class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape, self.dtype = tuple(shape), dtype
        self.low, self.high = np.full(self.shape, low, dtype=dtype), np.full(self.shape, high, dtype=dtype)


class Dict(dict):
    def __init__(self, spaces):
        super().__init__(spaces)


class Env:
    _np_random = None

    def reset(self, seed=None, options=None):
        if seed is not None:  # gymnasium.utils.seeding.np_random(seed) == Generator(PCG64(SeedSequence(seed)))
            self._np_random = np.random.default_rng(seed)

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random


gym = types.ModuleType("gymnasium")
gym.Env = Env
gym.spaces = types.ModuleType("gymnasium.spaces")
gym.spaces.Box, gym.spaces.Dict = Box, Dict
gym.envs = types.ModuleType("gymnasium.envs")
gym.envs.registration = types.ModuleType("gymnasium.envs.registration")
gym.envs.registration.register = lambda **kw: None
for name, mod in (("gymnasium", gym), ("gymnasium.spaces", gym.spaces), ("gymnasium.envs", gym.envs),
                  ("gymnasium.envs.registration", gym.envs.registration)):
    sys.modules[name] = mod


# ------------------------------------------------------------------ stand-in for mujoco
class _Named:
    def __init__(self, id_):
        self.id = id_
        self.pos = np.zeros(3)
        self.xpos = np.zeros(3)


class MjModel:
    @staticmethod
    def from_xml_path(path):
        m = MjModel()
        m.task = XML_TASK[os.path.basename(path)]
        m.c = mjcf.compile_model(os.path.dirname(path), m.task)
        c = m.c
        m.ncube = int(c["ncube"])
        m.nq, m.nv = 6 + 7 * m.ncube, 6 + 6 * m.ncube
        m.jnt_range = np.asarray(c["jnt_range"], dtype=np.float64).copy()
        m.actuator_ctrlrange = np.asarray(c["act_ctrlrange"], dtype=np.float64).copy()
        if "goal_center" in c:  # geom ids of goal_region_1 / goal_region_2 are 0 / 1 in this stand-in
            m.geom_pos = np.asarray(c["goal_center"], dtype=np.float64).copy()
            m.geom_size = np.stack([c["goal_size"], c["goal_size"]]).astype(np.float64)
        m._geoms = {}
        m.opt = types.SimpleNamespace(timestep=float(c["timestep"]))
        return m

    def body(self, name):
        return _Named(name)

    def site(self, name):
        return _Named(name)

    def geom(self, name):
        return self._geoms.setdefault(name, _Named(name))


CUBE_NAMES = {1: ("cube",), 2: ("cube_red", "cube_blue")}


class MjData:
    def __init__(self, model):
        c = model.c
        self.qpos = np.zeros(model.nq)
        for k in range(model.ncube):
            self.qpos[6 + 7 * k: 9 + 7 * k] = c["cube_pos0"][k]
            self.qpos[9 + 7 * k] = 1.0
        self.qvel = np.zeros(model.nv)
        self.ctrl = np.zeros(6)
        self.time = 0.0
        self._site = _Named("end_effector_site")
        self._bodies = {n: _Named(n) for n in CUBE_NAMES[model.ncube]}
        self._fk = None

    def site(self, id_):
        assert id_ == "end_effector_site"
        return self._site

    def body(self, id_):
        return self._bodies[id_]


def mj_forward(model, data):
    """Kinematics only (what the reference's glue reads back): independent numpy FK of the arm, cube poses from qpos."""
    c = model.c
    xpos, xmat, axis = mjcf.arm_kinematics(c, data.qpos[:6])
    sb = int(c["site_body"])
    data._site.xpos[:] = xpos[sb] + xmat[sb] @ c["site_pos"]
    for k, n in enumerate(CUBE_NAMES[model.ncube]):
        data._bodies[n].xpos[:] = data.qpos[6 + 7 * k: 9 + 7 * k]
    data._fk = (xpos, axis, sb)


def mj_step(model, data):
    raise AssertionError("the fixtures are generated with n_substeps=0: mj_step is never reached")


def mj_jacSite(model, data, jacp, jacr, site_id):
    xpos, axis, sb = data._fk
    jacp[:] = 0
    for j in range(6):
        if j < sb:  # joint j sits on arm body j + 1; the site's body is sb
            jacp[:, j] = np.cross(axis[j], data._site.xpos - xpos[j + 1])


mj = types.ModuleType("mujoco")
mj.MjModel, mj.MjData, mj.mj_forward, mj.mj_step, mj.mj_jacSite = MjModel, MjData, mj_forward, mj_step, mj_jacSite
mj.mj_name2id = lambda model, kind, name: {"goal_region_1": 0, "goal_region_2": 1}[name]
mj.mjtObj = types.SimpleNamespace(mjOBJ_GEOM=5)
mj.viewer = types.ModuleType("mujoco.viewer")
sys.modules["mujoco"], sys.modules["mujoco.viewer"] = mj, mj.viewer

sys.path.insert(0, REF)
import gym_lowcostrobot.envs as ref_envs  # noqa: E402  (the reference, unmodified)

CASES = [("reach", "ReachCubeEnv"), ("push", "PushCubeEnv"), ("lift", "LiftCubeEnv"), ("pick_place", "PickPlaceCubeEnv"),
         ("stack", "StackTwoCubesEnv"), ("push_loop", "PushCubeLoopEnv")]
K = 48


def sample_state(env, rng, i):
    """Random arm pose / velocities; the cube(s) placed so that about half of the samples are near the success boundary."""
    m, d = env.model, env.data
    task = m.task
    lo, hi = m.jnt_range[:6, 0], m.jnt_range[:6, 1]
    q = rng.uniform(0.9 * lo, 0.9 * hi)
    qpos = d.qpos.copy()
    qpos[:6] = q
    qvel = rng.uniform(-1, 1, m.nv)
    xpos, xmat, _ = mjcf.arm_kinematics(m.c, q)
    sb = int(m.c["site_body"])
    site = xpos[sb] + xmat[sb] @ m.c["site_pos"]
    near = i % 2 == 0
    r = 0.04 if near else 0.3

    def quat():
        v = rng.normal(size=4)
        return v / np.linalg.norm(v)

    if task in ("reach", "lift"):
        qpos[6:9] = site + rng.uniform(-r, r, 3)
    elif task in ("push", "pick_place"):
        qpos[6:9] = env.target_pos.astype(np.float64) + rng.uniform(-r, r, 3)
    elif task == "stack":
        qpos[6:9] = rng.uniform([-0.15, 0.05, 0.0], [0.15, 0.3, 0.05])
        qpos[13:16] = qpos[6:9] + [0, 0, 0.03] + rng.uniform(-r, r, 3)
        qpos[16:20] = quat()
    else:  # push_loop: over both goal regions and the strip between them
        qpos[6:9] = [rng.uniform(-0.09, 0.09), rng.uniform(0.10, 0.17), 0.015]
        if i % 4 == 0:  # well inside the current goal region (success -> goal switch)
            qpos[6:8] = [env.goal_region_1_center, env.goal_region_2_center][env.current_goal][:2] + rng.uniform(-0.0015, 0.0015, 2)
    qpos[9:13] = quat()
    return qpos, qvel


out_dir = os.path.join(ROOT, "tests", "golden", "ref_glue")
os.makedirs(out_dir, exist_ok=True)
import json  # noqa: E402

# default constructor arguments for every env x action mode, plus variants with non-default kwargs (the quirks they pin:
# LiftCube with block_gripper=True uses action[-1] = the 5th arm action for the gripper too, lift_cube_env.py:264;
# block_gripper=False adds an ignored action entry in Reach / PushCubeLoop; sampling ranges and thresholds)
RUNS = [(task, cls, mode, "", {}) for task, cls in CASES for mode in ("joint", "ee")] + [
    ("lift", "LiftCubeEnv", "joint", "blocked", dict(block_gripper=True)),
    ("reach", "ReachCubeEnv", "joint", "gripper", dict(block_gripper=False, distance_threshold=0.03, cube_xy_range=0.2)),
    ("push", "PushCubeEnv", "ee", "ranges", dict(cube_xy_range=0.2, target_xy_range=0.25, distance_threshold=0.08)),
    ("pick_place", "PickPlaceCubeEnv", "joint", "ranges", dict(goal_z_range=0.2, target_xy_range=0.1, block_gripper=True)),
    ("stack", "StackTwoCubesEnv", "ee", "ranges", dict(cube_xy_range=0.15, distance_threshold=0.02)),
    ("push_loop", "PushCubeLoopEnv", "joint", "gripper", dict(block_gripper=False)),
    ("push_loop", "PushCubeLoopEnv", "ee", "gripper", dict(block_gripper=False)),
]
for task, cls, mode, variant, extra in RUNS:
    if True:
        rng = np.random.default_rng(77)
        kw = dict(observation_mode="state", action_mode=mode, n_substeps=0, **extra)
        envs = {}
        has_reward_type = task not in ("lift", "push_loop")
        for rt in (("sparse", "dense") if has_reward_type else (None,)):
            envs[rt] = getattr(ref_envs, cls)(**kw, **({"reward_type": rt} if rt else {}))
        rec = {k: [] for k in ("seed", "goal_in", "obs0", "qpos", "qvel", "action", "obs", "reward", "terminated", "truncated",
                               "success", "ctrl", "qpos_after", "goal_after", "dense")}
        for i in range(K):
            rt = (("sparse", "dense")[i % 2]) if has_reward_type else None
            env = envs[rt]
            seed = 1000 + i
            goal_in = 0
            if task == "push_loop":
                goal_in = env.current_goal = (i // 2) % 2
            obs0, _ = env.reset(seed=seed)
            qpos, qvel = sample_state(env, rng, i)
            env.data.qpos[:] = qpos
            env.data.qvel[:] = qvel
            mj_forward(env.model, env.data)
            na = env.action_space.shape[0]
            action = rng.uniform(-1.2, 1.2, na).astype(np.float32)  # beyond +-1 sometimes: the clip is on the path
            obs, reward, terminated, truncated, info = env.step(action)
            flat = lambda o: np.concatenate([np.asarray(o[k], dtype=np.float32).ravel() for k in o])
            for k in obs:
                assert obs[k].dtype == np.float32, (k, obs[k].dtype)
            success = info.get("is_success", info.get("success", False))
            rec["seed"].append(seed); rec["goal_in"].append(goal_in); rec["obs0"].append(flat(obs0))
            rec["qpos"].append(qpos); rec["qvel"].append(qvel); rec["action"].append(action); rec["obs"].append(flat(obs))
            rec["reward"].append(float(reward)); rec["terminated"].append(bool(terminated)); rec["truncated"].append(bool(truncated))
            rec["success"].append(bool(success)); rec["ctrl"].append(np.asarray(env.data.ctrl, dtype=np.float64).copy())
            rec["qpos_after"].append(env.data.qpos.copy())
            rec["goal_after"].append(env.current_goal if task == "push_loop" else 0)
            rec["dense"].append(rt == "dense")
        path = os.path.join(out_dir, f"{task}_{mode}" + (f"__{variant}" if variant else "") + ".npz")
        np.savez_compressed(path, obs_keys=np.array(list(obs)), kwargs=np.array(json.dumps(extra)), **{k: np.asarray(v) for k, v in rec.items()})
        print(path, os.path.getsize(path), "success:", int(np.sum(rec["success"])), "of", K)
