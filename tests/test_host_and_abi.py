"""Host logic, model compiler and C-ABI surface (no GPU compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gym_lowcostrobot_b200 import capi, config, mjcf, model, spaces

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ASSETS = "/root/reference/gym_lowcostrobot/assets/low_cost_robot_6dof"


def test_library_exports_every_symbol_of_the_header():
    hdr = open(os.path.join(ROOT, "include", "lcrsim.h")).read()
    declared = sorted(set(re.findall(r"\b(lcr_[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(capi.SYMBOLS), set(declared) ^ set(capi.SYMBOLS)
    L = capi.lib()
    for sym in declared:
        assert hasattr(L, sym), sym
    assert L.lcr_sizeof_model() == C.sizeof(model.LcrModel)
    assert L.lcr_sizeof_cfg() == C.sizeof(model.LcrEnvCfg)
    assert b"lcrsim" in L.lcr_version()
    assert L.lcr_obs_dim(0) == 15 and L.lcr_obs_dim(1) == 18 and L.lcr_obs_dim(4) == 18


def test_action_and_obs_dims_follow_the_reference():
    for task, bg, joint, ee in (("reach", True, 5, 3), ("push", True, 5, 3), ("lift", False, 6, 4), ("stack", False, 6, 4)):
        assert config.BLOCK_GRIPPER_DEFAULT[task] == bg
        assert config.action_dim(config.make_cfg(task, action_mode="joint")) == joint
        assert config.action_dim(config.make_cfg(task, action_mode="ee")) == ee
        assert capi.lib().lcr_action_dim(C.byref(config.make_cfg(task, action_mode="ee"))) == ee
    cfg = config.make_cfg("pick_place", cube_xy_range=0.3, goal_z_range=0.1)
    np.testing.assert_allclose(cfg.cube_low[:], [-0.15, 0.015, 0.0])
    np.testing.assert_allclose(cfg.cube_high[:], [0.15, 0.25, 0.0])
    np.testing.assert_allclose(cfg.target_high[:], [0.15, 0.25, 0.1])
    with pytest.raises(ValueError):
        config.make_cfg("reach", action_mode="torque")


def test_create_fails_loudly_without_cuda():
    import torch

    import gym_lowcostrobot_b200 as glr

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(capi.LcrError):
        glr.make("ReachCube-v0", num_envs=2)
    with pytest.raises(capi.LcrError):
        glr.make("PushCubeLoop-v0", num_envs=2)
    with pytest.raises(KeyError):
        glr.make("NoSuchTask-v0")
    with pytest.raises(ValueError):
        glr.make("ReachCube-v0", observation_mode="depth")


def test_compiled_models_are_consistent():
    for task in mjcf.TASK_XML:
        m = model.load_compiled(task)
        s, verts = model.pack_model(m)
        assert s.nq == 6 + 7 * s.ncube and s.nv == 6 + 6 * s.ncube
        assert s.nmesh == 20 and s.npair == 120 and verts.shape == (s.nvert, 3)
        assert s.impratio == 100.0 and s.timestep == 0.002  # follower.xml <option> wins over the scene's
        assert list(s.geom_condim[17:23:2]) == [6, 6, 4][: 3]  # link_5_collision, link_6_collision, cube
        com = np.array(m["mesh_com"])
        assert np.all(np.abs(com - m["mesh_center"]) <= m["mesh_half"] + 1e-9)
    assert model.load_compiled("pick_place")["cube_mass"][0] == 10.0
    assert np.isclose(model.load_compiled("stack")["cube_inertia"][0, 0], 1.125e-5)


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference assets not mounted")
def test_compiled_assets_match_a_fresh_parse_of_the_reference_mjcf():
    for task in ("reach", "stack", "push_loop"):
        fresh = mjcf.compile_model(REF_ASSETS, task)
        stored = model.load_compiled(task)
        extra = ("wall_pos", "wall_size", "goal_center", "goal_size", "cube_pos0", "geom_solref") if task == "push_loop" else ()
        for k in extra + ("body_pos", "body_quat", "body_mass", "jnt_axis", "jnt_range", "verts", "mesh_vertadr", "pair_g1",
                  "geom_friction", "geom_solimp", "dof_invweight0", "body_invweight0", "meaninertia", "mesh_com"):
            np.testing.assert_allclose(np.asarray(fresh[k], float), np.asarray(stored[k], float), atol=1e-12, err_msg=k)


def test_spaces_shim():
    b = spaces.Box(-1.0, 1.0, (5,), np.float32)
    b.seed(0)
    x = b.sample()
    assert x.dtype == np.float32 and b.contains(x) and not b.contains(np.full(5, 2.0, np.float32))
    d = spaces.Dict({"a": b})
    assert d.contains(d.sample())
    bb = spaces.batch_space(b, 7)
    assert bb.shape == (7, 5)
