"""Analytic / independent known-answer tests that pin the float64 oracle.

The reference ships no numeric tests (tests/test_env.py:8-13 is gymnasium check_env only) and MuJoCo is not
installable here, so the oracle is pinned by closed-form results and by an independent numpy restatement of
the kinematics / dynamics (gym_lowcostrobot_b200.mjcf.arm_kinematics / arm_mass_matrix, Jacobian-sum form vs
the oracle's recursive forms).
"""
import numpy as np
import pytest

from gym_lowcostrobot_b200 import mjcf, model
from oracle.oracle import Oracle

H = 0.002


def test_pcg64_stream_is_numpy_bit_exact():
    o = Oracle("reach")
    for seed in (0, 1, 12345, 2**40 + 7):
        o.seed(seed)
        ref = np.random.default_rng(seed).random(16)
        got = np.array([o.rng_double() for _ in range(16)])
        assert np.array_equal(ref, got)


@pytest.mark.parametrize("task", ["reach", "push", "pick_place", "stack"])
def test_reset_draws_match_numpy_generator(task):
    # reference reset: np_random.uniform(low, high) for cube(s), then target (push_cube_env.py:312-320,
    # stack_two_cubes_env.py:312-315); gymnasium seeds np_random = default_rng(seed)
    o = Oracle(task)
    for seed in (0, 3, 99):
        obs = o.reset(seed=seed)
        g = np.random.default_rng(seed)
        lo, hi = np.array(o.cfg.cube_low[:]), np.array(o.cfg.cube_high[:])
        st = o.get_state()
        ncube = 2 if task == "stack" else 1
        for c in range(ncube):
            np.testing.assert_array_equal(st["qpos"][6 + 7 * c:9 + 7 * c], g.uniform(lo, hi))
            np.testing.assert_array_equal(st["qpos"][9 + 7 * c:13 + 7 * c], [1, 0, 0, 0])
        if task in ("push", "pick_place"):
            tgt = g.uniform(np.array(o.cfg.target_low[:]), np.array(o.cfg.target_high[:])).astype(np.float32)
            np.testing.assert_array_equal(obs[12:15], tgt)
        # second reset without seed continues the same stream
        o.reset()
        np.testing.assert_array_equal(o.get_state()["qpos"][6:9], g.uniform(lo, hi))
        assert np.all(st["qpos"][:6] == 0)


def test_fk_at_zero_pose():
    o = Oracle("reach")
    o.set_state(qpos=np.r_[np.zeros(6), 0, 0.2, 0.5, 1, 0, 0, 0])
    o.forward()
    np.testing.assert_allclose(o.get("site_xpos"), [0.002017, 0.212570, 0.168400], atol=5e-7)
    xpos = o.get("xpos").reshape(9, 3)
    ref = [[0, -0.012, 0.0409], [-0.0209, -0.012, 0.0563], [-0.0144, 0.0028, 0.1646], [-0.01435, 0.10328, 0.1673]]
    np.testing.assert_allclose(xpos[1:5], ref, atol=2e-5)
    axis = o.get("axis").reshape(6, 3)
    np.testing.assert_allclose(axis, [[0, 0, -1], [1, 0, 0], [-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 0, -1]], atol=1e-3)


def test_kinematics_and_mass_matrix_match_numpy_restatement():
    m = model.load_compiled("reach")
    rng = np.random.default_rng(0)
    o = Oracle("reach")
    for _ in range(10):
        q = rng.uniform(-1.5, 1.5, 6)
        o.set_state(qpos=np.r_[q, 0, 0.2, 0.5, 1, 0, 0, 0], qvel=np.zeros(12))
        o.forward()
        xpos, xmat, axis = mjcf.arm_kinematics(m, q)
        np.testing.assert_allclose(o.get("xpos").reshape(9, 3)[:7], xpos, atol=1e-13)
        np.testing.assert_allclose(o.get("xmat").reshape(9, 3, 3)[:7], xmat, atol=1e-13)
        M, _ = mjcf.arm_mass_matrix(m, q)
        np.testing.assert_allclose(o.get("M").reshape(12, 12)[:6, :6], M, atol=1e-13)
    M0, _ = mjcf.arm_mass_matrix(m, np.zeros(6), armature=False)
    np.testing.assert_allclose(np.diag(M0), [2.229e-3, 4.037e-3, 1.776e-3, 2.34e-4, 6e-6, 1.1e-5], rtol=2e-2, atol=6e-7)
    np.testing.assert_allclose(m["dof_invweight0"], [9.78192, 9.61595, 9.82961, 9.97732, 9.99935, 9.99895], rtol=1e-5)
    assert abs(m["body_mass"][1:].sum() - 0.226020) < 1e-6


def _numeric_bias(m, q, v, g=np.array([0, 0, -9.81])):
    """C(q, v) v + g(q) from the mass matrix by finite differences (independent of the oracle's RNE)."""
    eps = 1e-6
    n = 6
    dM = np.zeros((n, n, n))
    for k in range(n):
        dq = np.zeros(n)
        dq[k] = eps
        dM[:, :, k] = (mjcf.arm_mass_matrix(m, q + dq, False)[0] - mjcf.arm_mass_matrix(m, q - dq, False)[0]) / (2 * eps)
    c = np.einsum("ijk,j,k->i", dM, v, v) - 0.5 * np.einsum("jki,j,k->i", dM, v, v)

    def pot(qq):
        xpos, xmat, _ = mjcf.arm_kinematics(m, qq)
        return -sum(m["body_mass"][b] * g @ (xpos[b] + xmat[b] @ m["body_ipos"][b]) for b in range(1, 7))

    grav = np.array([(pot(q + eps * np.eye(n)[k]) - pot(q - eps * np.eye(n)[k])) / (2 * eps) for k in range(n)])
    return c + grav


def test_bias_forces_match_lagrangian_finite_differences():
    m = model.load_compiled("reach")
    rng = np.random.default_rng(1)
    o = Oracle("reach")
    o.set_state(qpos=np.r_[np.zeros(6), 0, 0.2, 0.5, 1, 0, 0, 0], qvel=np.zeros(12))
    o.forward()
    np.testing.assert_allclose(o.get("bias")[:6], [0, 0.147385, -0.129713, 0.034035, 0.000725, 0], atol=2e-6)
    for _ in range(5):
        q, v = rng.uniform(-1.2, 1.2, 6), rng.uniform(-3, 3, 6)
        o.set_state(qpos=np.r_[q, 0, 0.2, 0.5, 1, 0, 0, 0], qvel=np.r_[v, np.zeros(6)])
        o.forward()
        np.testing.assert_allclose(o.get("bias")[:6], _numeric_bias(m, q, v), atol=2e-8)
    np.testing.assert_allclose(o.get("bias")[6:9], [0, 0, 0.1 * 9.81])


def test_free_fall_is_semi_implicit_euler():
    # z_k = z0 - 1/2 g h^2 k (k + 1) exactly until contact
    o = Oracle("reach")
    z0 = 0.5
    o.set_state(qpos=np.r_[np.zeros(6), 0.1, 0.2, z0, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=np.zeros(6))
    for k in range(1, 60):
        o.substep()
        z = o.get_state()["qpos"][8]
        assert abs(z - (z0 - 0.5 * 9.81 * H * H * k * (k + 1))) < 1e-12
        assert o.diag()["ncon"] == 0 or k > 1


@pytest.mark.parametrize("task", ["reach", "pick_place"])  # cube mass 0.1 and 10
def test_cube_rest_height_is_mass_independent(task):
    o = Oracle(task)
    o.set_state(qpos=np.r_[np.zeros(6), 0.1, 0.2, 0.0149, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=np.zeros(6))
    o.substep(1500)
    z = o.get_state()["qpos"][8]
    assert abs(z - 0.0148922) < 2e-6, z
    assert o.diag()["ncon"] == 4 and o.diag()["nefc"] == 16
    con = o.get("contacts").reshape(-1, 27)
    assert np.all(con[:, 13] == 4) and np.allclose(con[:, 17], 0.5)  # condim 4, cube friction wins by priority
    np.testing.assert_allclose(con[:, 12], z - 0.015, atol=1e-9)


def test_contact_parameter_classes():
    o = Oracle("lift")
    # gripper finger (link_6_collision, geom 19) pressed into the floor
    q = np.array([0, 1.2, 1.0, 1.0, 0, 0.0])
    o.set_state(qpos=np.r_[q, 0.3, 0.3, 0.0149, 1, 0, 0, 0], qvel=np.zeros(12))
    o.forward()
    con = o.get("contacts").reshape(-1, 27)
    geoms = {(int(c[14]), int(c[15])): c for c in con}
    finger = [c for (g1, g2), c in geoms.items() if g2 in (17, 19)]
    other = [c for (g1, g2), c in geoms.items() if g1 == 20 and g2 < 17]
    if finger:
        c = finger[0]
        assert c[13] == 6 and np.isclose(c[17], 1.5) and np.isclose(c[22], 0.015) and np.isclose(c[24], 0.036)
    for c in other:
        assert c[13] == 3 and np.isclose(c[17], 1.0)
    # default K, B of solref (0.02, 1) with dmax 0.95
    assert np.isclose(1 / (0.95**2 * 0.02**2), 2770.083, rtol=1e-6) and np.isclose(2 / (0.95 * 0.02), 105.2632, rtol=1e-6)


def test_servo_step_matches_closed_form_on_joint_1():
    # joint_1 has a vertical axis: no gravity torque; M11 + armature + h (damping + kv) in the implicit update
    o = Oracle("reach")
    m = model.load_compiled("reach")
    o.set_state(qpos=np.r_[np.zeros(6), 0.3, 0.3, 0.0149, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=np.r_[1.0, np.zeros(5)])
    o.forward()
    assert np.isclose(o.get("actuator")[0], 10.0)  # 1000 * 1 rad clamps at +10 N m
    M = o.get("M").reshape(12, 12)[:6, :6]
    smooth = o.get("actuator")[:6] + o.get("passive")[:6] - o.get("bias")[:6]
    qacc = np.linalg.solve(M, smooth)
    a = np.linalg.solve(M + H * np.diag(m["jnt_damping"] + m["act_kv"]), M @ qacc)
    o.set_state(qpos=np.r_[np.zeros(6), 0.3, 0.3, 0.0149, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=np.r_[1.0, np.zeros(5)])
    o.substep()
    st = o.get_state()
    np.testing.assert_allclose(st["qvel"][:6], H * a, atol=1e-10)
    np.testing.assert_allclose(st["qpos"][:6], H * H * a, atol=1e-12)
    assert 60 < a[0] < 100  # |qdd| ~ 10 / (0.1 + M11 + 0.022)


def test_joint_action_map_and_limits():
    o = Oracle("lift")
    o.reset(seed=0)
    a = np.array([2.0, -2.0, 0.5, 0.25, -0.75, 1.0], np.float32)  # clipped to [-1, 1] first
    o.step(a)
    lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
    hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])
    np.testing.assert_allclose(o.get_state()["ctrl"], np.clip(np.clip(a, -1, 1), lo, hi), atol=1e-12)
    r = Oracle("reach")
    r.reset(seed=0)
    r.step(np.ones(5, np.float32))
    assert r.get_state()["ctrl"][5] == 0.0  # blocked gripper (reach_cube_env.py:255)


def test_ik_properties():
    o = Oracle("reach", action_mode="ee")
    o.reset(seed=0)
    site = o.get_state()["aux"][4:7]
    q0 = o.get_state()["qpos"][:6].copy()
    q = o.ik(site + np.array([0.004, 0, 0]))  # within tolerance 0.01: zero iterations
    np.testing.assert_allclose(q, q0, atol=1e-7)
    q = o.ik(site + np.array([0.0, -0.05, 0.03]))
    assert np.abs(q - q0).max() <= 10 * 0.5 + 1e-6 and q[5] == q0[5]  # gripper column of the Jacobian is zero
    np.testing.assert_allclose(o.get_state()["qpos"][:6], q0)  # lcr_ik-style call leaves the state untouched
    # in-step IK teleports the arm (reach_cube_env.py:185): qpos jumps by more than one substep could move it
    o.step(np.array([0.0, -1.0, 1.0], np.float32))
    assert np.abs(o.get_state()["ctrl"][:5] - q0[:5]).max() > 0.05


def test_time_limit_and_termination():
    o = Oracle("lift")
    o.reset(seed=1)
    for k in range(50):
        obs, r, te, tr, su = o.step(np.zeros(6, np.float32))
        assert te is False and tr == (k == 49)
    p = Oracle("push", distance_threshold=10.0)
    p.reset(seed=1)
    obs, r, te, tr, su = p.step(np.zeros(5, np.float32))
    assert te and su and r == 0.0
    d = Oracle("push", reward_type="dense")
    d.reset(seed=1)
    obs, r, te, tr, su = d.step(np.zeros(5, np.float32))
    assert np.isclose(-r, np.linalg.norm(obs[15:18].astype(np.float64) - obs[12:15]), atol=2e-3)


def test_box_box_stack_is_stable():
    o = Oracle("stack")
    qpos = np.r_[np.zeros(6), 0.1, 0.2, 0.0149, 1, 0, 0, 0, 0.1, 0.2, 0.0448, 1, 0, 0, 0]
    o.set_state(qpos=qpos, qvel=np.zeros(18), ctrl=np.zeros(6))
    o.substep(1000)
    st = o.get_state()
    assert abs(st["qpos"][15] - 0.0447) < 3e-4 and abs(st["qpos"][8] - 0.01489) < 1e-4
    assert np.abs(st["qvel"][6:]).max() < 1e-6
    assert o.diag()["ncon"] == 8


def test_mpr_box_mesh_and_self_collision_contacts_are_sane():
    o = Oracle("push")
    # cube pushed 5 mm into the side of link_5 (site is on link_5)
    o.set_state(qpos=np.r_[np.zeros(6), 0.0, 0.2, 0.5, 1, 0, 0, 0])
    o.forward()
    site = o.get("site_xpos")
    o.set_state(qpos=np.r_[np.zeros(6), site[0], site[1] + 0.012, site[2], 1, 0, 0, 0])
    o.forward()
    con = o.get("contacts").reshape(-1, 27)
    cm = con[con[:, 14] == 21]
    assert len(cm) >= 1
    for c in cm:
        assert -0.03 < c[12] < 0 and abs(np.linalg.norm(c[3:6]) - 1) < 1e-9
        assert c[4] < -0.5  # normal from the cube (geom1) towards the mesh (geom2): -y here
    assert o.diag()["overflow"] == 0


# ------------------------------------------------------------------ PushCubeLoop scene (push_cube_loop.xml, push_cube_loop_env.py)
def test_push_loop_model_constants():
    m = model.load_compiled("push_loop")
    s, _ = model.pack_model(m)
    assert s.task == 5 and s.ncube == 1 and s.nwall == 4
    np.testing.assert_allclose(np.ctypeslib.as_array(s.wall_pos), [[-0.125, 0.135, 0.005], [0.125, 0.135, 0.005], [0, 0.09, 0.005], [0, 0.18, 0.005]])
    np.testing.assert_allclose(np.ctypeslib.as_array(s.wall_size), [[0.01, 0.055, 0.007]] * 2 + [[0.125, 0.01, 0.007]] * 2)
    np.testing.assert_allclose(np.ctypeslib.as_array(s.goal_center), [[0.06, 0.135, 0.01], [-0.06, 0.135, 0.01]])
    np.testing.assert_allclose(np.ctypeslib.as_array(s.cube_pos0)[0], [0.06, 0.135, 0.017])
    assert s.cube_mass[0] == 0.05 and s.impratio == 100.0
    g = s.nmesh  # floor, cube, walls
    assert list(np.ctypeslib.as_array(s.geom_solref)[g]) == [0.0, 0.0]  # floor solref="0 0" (push_cube_loop.xml:25)
    assert s.geom_condim[g + 1] == 4 and s.geom_priority[g + 1] == 1
    np.testing.assert_allclose(np.ctypeslib.as_array(s.geom_friction)[g + 1], [1.5, 1.5, 1.5])
    assert list(s.geom_condim[g + 2:g + 6]) == [3, 3, 3, 3]


def test_push_loop_reset_samples_inside_the_current_goal_region():
    """push_cube_loop_env.py:302-320 with numpy's generator: low/high = (+-0.0095, +-0.0145, 0.0035), shifted to the
    centre of the current goal region; the goal persists across resets."""
    hi = np.array([0.035, 0.045, 0.007]) / 2
    hi[:2] -= 0.008
    lo = hi * np.array([-1.0, -1.0, 1.0])
    centers = np.array([[0.06, 0.135], [-0.06, 0.135]])
    for goal in (0, 1):
        o = Oracle("push_loop")
        st = o.get_state()
        st["aux"][1] = goal
        o.set_state(aux=st["aux"])
        for seed in (0, 5, 99):
            obs = o.reset(seed=seed)
            p = np.random.default_rng(seed).uniform(lo, hi)
            p[:2] += centers[goal]
            np.testing.assert_array_equal(obs[12:15], p.astype(np.float32))
            np.testing.assert_array_equal(obs[:12], 0)


def _loop_reward_numpy(cube_xy, goal):
    """get_reward / get_cube_overlap (push_cube_loop_env.py:337-383) restated with the reference's dtypes."""
    cube = np.asarray(cube_xy, dtype=np.float64).astype(np.float32)
    centers = np.array([[0.06, 0.135, 0.01], [-0.06, 0.135, 0.01]])
    high = np.array([0.035, 0.045, 0.007]) / 2
    high[:2] -= 0.008
    low = high * np.array([-1.0, -1.0, 1.0])
    w = 0.015 / 2
    gx, gy = centers[goal][:2]
    wg, lg = high[:2]
    xo = max(0, min(cube[0] + w, gx + wg) - max(cube[0] - w, gx - wg))
    yo = max(0, min(cube[1] + w, gy + lg) - max(cube[1] - w, gy - lg))
    overlap = xo * yo / (w * w * 4)
    if overlap > 0.95:
        return 5.0, 1, 1 - goal
    if overlap > 0.0:
        return float(overlap - 1), 0, goal
    edge = low[1] + centers[goal][1]
    d = np.sqrt((cube[1] - edge) ** 2)
    return float(min(max((-d / 0.16) - 1, -2), -1)), 0, goal


def test_push_loop_reward_overlap_and_goal_switch():
    rng = np.random.default_rng(11)
    o = Oracle("push_loop", n_substeps=0, max_episode_steps=0)
    seen = set()
    for k in range(400):
        goal = k % 2
        xy = np.array([rng.uniform(-0.09, 0.09), rng.uniform(0.1, 0.17)])
        if k % 5 == 0:
            xy = np.array([0.06, 0.135]) * [1 - 2 * goal, 1] + rng.uniform(-0.002, 0.002, 2)
        st = o.get_state()
        st["aux"][1] = goal
        qpos = st["qpos"].copy()
        qpos[6:8] = xy
        o.set_state(qpos=qpos, aux=st["aux"])
        obs, r, te, tr, su = o.step(np.zeros(5, np.float32))
        r_ref, su_ref, goal_ref = _loop_reward_numpy(xy, goal)
        assert r == np.float32(r_ref) and su == bool(su_ref) and not te and not tr, (k, r, r_ref)
        assert int(o.get_state()["aux"][1]) == goal_ref
        seen.add("success" if su else ("partial" if r > -1 else "outside"))
    assert seen == {"success", "partial", "outside"}


def test_push_loop_time_limit_truncates_and_never_terminates():
    o = Oracle("push_loop")
    o.reset(seed=0)
    flags = [o.step(np.zeros(5, np.float32))[2:4] for _ in range(50)]
    assert not any(f[0] for f in flags) and [f[1] for f in flags] == [False] * 49 + [True]


def test_push_loop_rails_contain_the_cube():
    """A cube sliding at 0.6 m/s into each rail from 4 mm away (floor friction 1.5 would stop it within 12 mm) is stopped
    by the rail: it never gets further than the rail's inner face (plus a transient soft-contact penetration) and stays
    on the floor inside the 0.23 x 0.07 m pen."""
    # (axis, direction) -> coordinate of the cube centre when its face touches the rail's inner face
    inner = {(0, 1): 0.115 - 0.015, (0, -1): -(0.115 - 0.015), (1, 1): 0.17 - 0.015, (1, -1): 0.10 + 0.015}
    for (ax, sg), lim in inner.items():
        o = Oracle("push_loop")
        xy = np.array([0.0, 0.135])
        xy[ax] = lim - 0.004 * sg
        qpos = np.r_[np.zeros(6), xy, 0.0149, 1, 0, 0, 0]
        qvel = np.zeros(12)
        qvel[6 + ax] = 0.6 * sg
        o.set_state(qpos=qpos, qvel=qvel, ctrl=np.zeros(6))
        far, wall_contacts = -1.0, 0
        for _ in range(300):
            o.substep(1)
            far = max(far, sg * (o.get_state()["qpos"][6 + ax] - lim))
            con = o.get("contacts").reshape(-1, 27)
            wall_contacts += int(np.sum((con[:, 14] >= 22) & (con[:, 15] == 21)))  # geom1 = a wall (geoms 22..25), geom2 = the cube
        assert wall_contacts > 0, (ax, sg)
        # soft contact (solref time constant 0.02 s): the 0.6 m/s impact penetrates a few mm, then the cube is pushed back out
        assert -1e-3 < far < 4e-3, (ax, sg, far)
        st = o.get_state()
        assert sg * (st["qpos"][6 + ax] - lim) < 2e-4, (ax, sg, st["qpos"][6:8])
        assert abs(st["qpos"][8] - 0.01489) < 5e-4 and np.abs(st["qvel"][6:9]).max() < 0.05


def test_push_loop_arm_collides_with_the_rails():
    """Gripper driven (by the IK helper) onto the two long rails: wall-mesh contacts appear with the wall as geom1 and a
    moving arm mesh as geom2, unit normals pointing from the rail into the arm (upwards for a touch from above)."""
    o = Oracle("push_loop", collision_mask=model.COLLIDE_WALL_MESH)
    found, up = 0, 0
    rng = np.random.default_rng(3)
    for k in range(40):
        target = np.array([rng.uniform(-0.1, 0.1), (0.09, 0.18)[k % 2], rng.uniform(0.0, 0.015)])
        q = np.zeros(6)
        for _ in range(4):  # 4 x 10 damped-least-squares iterations
            o.set_state(qpos=np.r_[q, 0.06, 0.135, 0.0149, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=q)
            o.forward()
            q = o.ik(target).astype(np.float64)
        o.set_state(qpos=np.r_[q, 0.06, 0.135, 0.0149, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=q)
        o.forward()
        for c in o.get("contacts").reshape(-1, 27):
            assert c[14] >= 22 and c[15] < 20  # wall first, arm mesh second
            assert c[12] < 0 and abs(np.linalg.norm(c[3:6]) - 1) < 1e-9
            found += 1
            up += c[5] > 0.5
    assert found >= 20 and up >= found // 2, (found, up)


@pytest.mark.parametrize("task", ["push", "stack", "push_loop"])
def test_separating_axis_cache_never_changes_the_physics(task):
    """The cache of the convex narrowphase only decides HOW a separation is proven (cached axis with vertex-free bounds,
    exact support test, or MPR), never the contacts: a rollout whose cache is emptied before every substep (set_state does
    that) must be bit-identical to the free-running one."""
    import ctypes

    from oracle.oracle import lib

    stats = (ctypes.c_long * 8)()
    lib().orc_stats(stats, 1)
    rng = np.random.default_rng(5)
    for seed in range(3):
        a, b = Oracle(task), Oracle(task)
        a.reset(seed=seed)
        b.reset(seed=seed)
        lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706])
        hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706])
        ncon = 0
        for step in range(4):
            ctrl = np.r_[np.clip(a.get_state()["qpos"][:5] + rng.uniform(-1, 1, 5), lo, hi), 0.0]
            for o in (a, b):
                o.set_state(ctrl=ctrl)
            for k in range(20):
                a.substep(1)
                st = b.get_state()
                b.set_state(qpos=st["qpos"], qvel=st["qvel"], ctrl=st["ctrl"], warm=st["warm"])  # empties b's cache
                b.substep(1)
                ncon += a.diag()["ncon"]
            sa, sb = a.get_state(), b.get_state()
            np.testing.assert_array_equal(sa["qpos"], sb["qpos"])
            np.testing.assert_array_equal(sa["qvel"], sb["qvel"])
        assert ncon > 0
    lib().orc_stats(stats, 0)
    assert stats[0] > 0 and stats[6] > 0, list(stats)  # MPR ran, and cached axes answered broadphase survivors


def _constraint_cost(a, M, a0, J, aref, D, nlim, contacts):
    """MuJoCo's primal objective, restated in numpy from the published formulation (Gauss term + limit rows + elliptic
    cones in their three zones): 1/2 (a - a0)^T M (a - a0) + sum_units s(J a - aref)."""
    jar = J @ a - aref
    da = a - a0
    cost = 0.5 * da @ M @ da
    for i in range(nlim):
        if jar[i] < 0:
            cost += 0.5 * D[i] * jar[i] ** 2
    i = nlim
    for dim, mu, fr in contacts:
        x = jar[i:i + dim]
        if dim == 1:
            cost += 0.5 * D[i] * x[0] ** 2 if x[0] < 0 else 0.0
        else:
            scale = np.r_[mu, fr[:dim - 1]]
            u = x * scale
            N, T = u[0], np.linalg.norm(u[1:])
            if N >= mu * T or (T <= 0 and N >= 0):
                pass
            elif mu * N + T <= 0 or (T <= 0 and N < 0):
                cost += 0.5 * np.sum(D[i:i + dim] * x * x)
            else:
                cost += 0.5 * D[i] / (mu * mu * (1 + mu * mu)) * (N - mu * T) ** 2
        i += dim
    return cost


@pytest.mark.parametrize("task", ["push", "stack", "push_loop"])
def test_newton_solution_minimises_the_published_objective(task):
    """The oracle's qacc must be a minimiser of MuJoCo's convex primal objective evaluated by an independent numpy
    restatement: a general-purpose optimiser (scipy BFGS, restarted from the oracle's point and from qacc_smooth) finds
    no point that is better by more than the solver tolerance."""
    from scipy.optimize import minimize

    rng = np.random.default_rng(21)
    m = model.load_compiled(task)
    lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
    hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])
    checked, zones = 0, 0
    for trial in range(40):
        o = Oracle(task)
        nq, nv = o.nq, o.nv
        qpos = np.zeros(nq)
        qpos[:6] = rng.uniform(lo, hi)
        xpos, _, _ = mjcf.arm_kinematics(m, qpos[:6])
        for c in range((nq - 6) // 7):
            p = qpos[6 + 7 * c: 13 + 7 * c]
            p[:3] = xpos[rng.integers(1, 7)] + rng.uniform(-0.03, 0.03, 3) if trial % 2 else [rng.uniform(-0.1, 0.1), rng.uniform(0.1, 0.2), rng.uniform(0.0, 0.02)]
            q = rng.normal(size=4)
            p[3:] = q / np.linalg.norm(q)
        o.set_state(qpos=qpos, qvel=rng.normal(scale=0.5, size=nv), ctrl=rng.uniform(lo, hi), warm=np.zeros(nv))
        o.forward()
        d = o.diag()
        if d["nefc"] == 0 or d["overflow"]:
            continue
        nefc = d["nefc"]
        M = o.get("M").reshape(nv, nv)
        J = o.get("efc_J").reshape(nefc, nv)
        D, aref, a0, a = o.get("efc_D"), o.get("efc_aref"), o.get("qacc_smooth"), o.get("qacc")
        con = o.get("contacts").reshape(-1, 27)
        contacts = [(int(c[13]), c[16], c[17:22]) for c in con]
        nlim = nefc - sum(c[0] for c in contacts)
        f = lambda x: _constraint_cost(x, M, a0, J, aref, D, nlim, contacts)
        c_oracle = f(a)
        best = min(minimize(f, x0, method="BFGS", options={"gtol": 1e-10, "maxiter": 2000}).fun for x0 in (a, a0))
        scale = abs(c_oracle) + 1e-3
        assert c_oracle <= best + 1e-6 * scale, (task, trial, c_oracle, best, d)
        assert c_oracle <= f(a0) + 1e-12  # never worse than the unconstrained acceleration
        checked += 1
        zones += c_oracle > 1e-9
    assert checked >= 15 and zones >= 8, (checked, zones)


@pytest.mark.parametrize("task", ["push", "stack", "push_loop"])
def test_contact_jacobian_matches_finite_differences_of_the_kinematics(task):
    """Row r of efc_J times a generalized velocity v is the relative velocity of the two bodies' material points at the
    contact (translational rows) / their relative angular velocity (torsional and rolling rows) along the contact frame
    axes, body 2 minus body 1.  Checked against central differences of the forward kinematics alone (cube angular
    velocity is body-local, as in MuJoCo's free joint)."""
    rng = np.random.default_rng(8)
    m = model.load_compiled(task)
    nmesh = len(m["mesh_body"])
    lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
    hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])

    def body_of(g, ncube):
        if g < nmesh:
            return int(m["mesh_body"][g])
        c = g - nmesh - 1
        return 7 + c if 0 <= c < ncube else -1  # floor / static walls: the world

    def integrate(qpos, v, eps, ncube):
        q = qpos.copy()
        q[:6] += eps * v[:6]
        for c in range(ncube):
            q[6 + 7 * c: 9 + 7 * c] += eps * v[6 + 6 * c: 9 + 6 * c]
            w = eps * v[9 + 6 * c: 12 + 6 * c]  # body-local rotation vector
            ang = np.linalg.norm(w)
            dq = np.r_[np.cos(ang / 2), (np.sin(ang / 2) / ang) * w] if ang > 0 else np.array([1.0, 0, 0, 0])
            q[9 + 7 * c: 13 + 7 * c] = mjcf.quat_mul(q[9 + 7 * c: 13 + 7 * c], dq)
        return q

    def poses(o, qpos, nv):
        o.set_state(qpos=qpos, qvel=np.zeros(nv))
        o.forward()
        return o.get("xpos").reshape(9, 3).copy(), o.get("xmat").reshape(9, 3, 3).copy()

    rows_checked = 0
    for trial in range(12):
        o = Oracle(task)
        nq, nv = o.nq, o.nv
        ncube = (nq - 6) // 7
        qpos = np.zeros(nq)
        qpos[:6] = rng.uniform(lo, hi)
        xp, _, _ = mjcf.arm_kinematics(m, qpos[:6])
        for c in range(ncube):
            p = qpos[6 + 7 * c: 13 + 7 * c]
            p[:3] = xp[rng.integers(1, 7)] + rng.uniform(-0.03, 0.03, 3) if trial % 2 else [rng.uniform(-0.1, 0.1), rng.uniform(0.1, 0.2), 0.01]
            q = rng.normal(size=4)
            p[3:] = q / np.linalg.norm(q)
        x0, R0 = poses(o, qpos, nv)
        d = o.diag()
        if d["nefc"] == 0 or d["overflow"]:
            continue
        J = o.get("efc_J").reshape(d["nefc"], nv)
        con = o.get("contacts").reshape(-1, 27)
        nlim = d["nefc"] - int(con[:, 13].sum())
        v = rng.normal(size=nv)
        eps = 1e-6
        (xa, Ra), (xb, Rb) = poses(o, integrate(qpos, v, eps, ncube), nv), poses(o, integrate(qpos, v, -eps, ncube), nv)
        row = nlim
        for c in con:
            dim, frame, pos = int(c[13]), c[3:12].reshape(3, 3), c[0:3]
            rel_v, rel_w = np.zeros(3), np.zeros(3)
            for b, sg in ((body_of(int(c[14]), ncube), -1.0), (body_of(int(c[15]), ncube), 1.0)):
                if b < 0:
                    continue
                local = R0[b].T @ (pos - x0[b])
                rel_v += sg * ((xa[b] + Ra[b] @ local) - (xb[b] + Rb[b] @ local)) / (2 * eps)
                W = (Ra[b] @ Rb[b].T - Rb[b] @ Ra[b].T) / (4 * eps)  # skew part of dR R^T / (2 eps)
                rel_w += sg * np.array([W[2, 1], W[0, 2], W[1, 0]])
            expect = np.r_[frame @ rel_v, frame @ rel_w][:dim if dim <= 3 else None]
            if dim == 1:
                expect = expect[:1]
            elif dim == 4:
                expect = np.r_[frame @ rel_v, (frame @ rel_w)[:1]]
            np.testing.assert_allclose(J[row:row + dim] @ v, expect[:dim], rtol=0, atol=2e-7)
            row += dim
            rows_checked += dim
    assert rows_checked >= 60, rows_checked


@pytest.mark.parametrize("task,mu,mass", [("push", 0.5, 0.1), ("pick_place", 0.5, 10.0), ("push_loop", 1.5, 0.05)])
def test_sliding_contacts_sit_on_the_coulomb_cone(task, mu, mass):
    """A cube sliding flat on the floor: every active contact carries a friction force of exactly mu times its normal
    force, opposed to the sliding direction (elliptic cone boundary; mu = the cube geom's sliding friction, which has
    priority over the floor's: push_cube.xml:28, push_cube_loop.xml:31), the four normal forces carry the weight and the
    normal acceleration the solver asks for, and the cube's deceleration is the friction sum over its mass."""
    o = Oracle(task)
    assert np.isclose(model.load_compiled(task)["cube_mass"][0], mass)
    qpos = np.r_[np.zeros(6), 0.0, 0.135, 0.014892, 1, 0, 0, 0]
    qvel = np.zeros(12)
    qvel[6] = 0.4
    o.set_state(qpos=qpos, qvel=qvel, ctrl=np.zeros(6))
    o.substep(1)
    con = o.get("contacts").reshape(-1, 27)
    f = o.get("efc_force")
    assert len(con) == 4 and all(int(c[13]) == 4 for c in con)
    fx, active = 0.0, 0
    for k, c in enumerate(con):
        fn, ft = f[4 * k], f[4 * k + 1: 4 * k + 3]
        frame = c[3:12].reshape(3, 3)
        if fn == 0:  # mu > 1: the friction torque unloads the trailing edge (the cube starts to tip)
            assert mu > 1 and np.all(ft == 0)
            continue
        active += 1
        np.testing.assert_allclose(np.linalg.norm(ft), mu * fn, rtol=1e-6)
        world = fn * frame[0] + ft[0] * frame[1] + ft[1] * frame[2]  # force on the cube (geom 2 of the floor-cube pair)
        assert world[0] < 0 and abs(world[1]) < 1e-9 * fn + 1e-12
        fx += world[0]
        assert abs(f[4 * k + 3]) < 1e-9  # no torsion
    assert active == (4 if mu < 1 else 2)
    decel = (0.4 - o.get_state()["qvel"][6]) / 0.002
    np.testing.assert_allclose(decel, -fx / mass, rtol=1e-9)


@pytest.mark.parametrize("task,mass", [("push", 0.1), ("pick_place", 10.0), ("push_loop", 0.05)])
def test_resting_cube_is_carried_by_its_weight(task, mass):
    """Equilibrium on the floor: the four corner contacts share the weight equally and carry no friction."""
    o = Oracle(task)
    o.set_state(qpos=np.r_[np.zeros(6), 0.0, 0.135, 0.0149, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=np.zeros(6))
    o.substep(600)
    f = o.get("efc_force").reshape(4, 4)
    np.testing.assert_allclose(f[:, 0], mass * 9.81 / 4, rtol=1e-6)
    assert np.abs(f[:, 1:]).max() < 1e-9 * mass * 9.81 + 1e-12
    assert np.abs(o.get_state()["qvel"][6:]).max() < 1e-9


def test_joint_limit_row_matches_the_published_impedance_formulas():
    """Joint 3 held 5 mm-equivalent (0.005 rad) beyond its upper limit at rest: one limit row with
    pos = -r, impedance d = dmax = 0.95 (|pos| >= width 0.001), aref = -B vel - K d pos with K = 1 / (dmax^2 tc^2 zeta^2),
    B = 2 / (dmax tc) for solref (0.02, 1), R = (1 - d) / d * dof_invweight0, D = 1 / R, Jacobian -e_3."""
    o = Oracle("reach", collision_mask=0)
    m = model.load_compiled("reach")
    r, j = 0.005, 2
    q = np.zeros(6)
    q[j] = m["jnt_range"][j][1] + r
    for vel in (0.0, 0.3):
        qvel = np.zeros(12)
        qvel[j] = vel
        o.set_state(qpos=np.r_[q, 0.0, 0.3, 0.0149, 1, 0, 0, 0], qvel=qvel, ctrl=q)
        o.forward()
        assert o.diag()["nefc"] == 1
        J = o.get("efc_J").reshape(1, 12)
        expect_J = np.zeros(12)
        expect_J[j] = -1.0
        np.testing.assert_array_equal(J[0], expect_J)
        np.testing.assert_allclose(o.get("efc_pos")[0], -r, atol=1e-15)
        K, B, d = 1 / (0.95**2 * 0.02**2), 2 / (0.95 * 0.02), 0.95
        np.testing.assert_allclose(o.get("efc_aref")[0], -B * (-vel) - K * d * (-r), rtol=1e-12)
        R = (1 - d) / d * m["dof_invweight0"][j]
        np.testing.assert_allclose(o.get("efc_R")[0], R, rtol=1e-12)
        np.testing.assert_allclose(o.get("efc_D")[0], 1 / R, rtol=1e-12)
        f = o.get("efc_force")[0]
        assert f >= 0  # unilateral: can only push the joint back inside (0 when the servo, whose target is clamped to the range, already does)
        np.testing.assert_allclose(o.get("qfrc_constraint")[j], -f, rtol=1e-12, atol=1e-15)
        jar = J[0] @ o.get("qacc") - o.get("efc_aref")[0]
        np.testing.assert_allclose(f, max(0.0, -jar / R), rtol=1e-9, atol=1e-12)  # f = -D min(0, J a - aref)


@pytest.mark.parametrize("task,mass,fr", [("push", 0.1, (0.5, 0.005)), ("push_loop", 0.05, (1.5, 1.5))])
def test_contact_rows_match_the_published_regularisation(task, mass, fr):
    """Floor-cube corner contact (condim 4, cube priority 1) of a cube at rest, penetration r at the corners:
    normal row R0 = (1 - d) / d * (1 / m) with d = d(r) from solimp (0.9, 0.95, 0.001, 0.5, 2), aref0 = -B v - K d (-r);
    elliptic cone: R1 = R2 = R0 / impratio, R3 = R1 mu1^2 / mu_torsion^2, friction rows have no position term, and the
    regularised cone's mu = mu1 sqrt(R1 / R0) = mu1 / sqrt(impratio) (impratio = 100 from follower.xml:3)."""
    o = Oracle(task)
    z = 0.0148
    o.set_state(qpos=np.r_[np.zeros(6), 0.0, 0.135, z, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=np.zeros(6))
    o.forward()
    con = o.get("contacts").reshape(-1, 27)
    assert len(con) == 4 and o.diag()["nefc"] == 16
    r = 0.015 - z
    x = r / 0.001
    assert x < 1
    y = x**2 / 0.5 if x <= 0.5 else 1 - (1 - x) ** 2 / 0.5  # power 2, midpoint 0.5
    d = 0.9 + y * 0.05
    K, B = 1 / (0.95**2 * 0.02**2), 2 / (0.95 * 0.02)
    R, aref, pos = o.get("efc_R"), o.get("efc_aref"), o.get("efc_pos")
    for k, c in enumerate(con):
        assert int(c[13]) == 4
        np.testing.assert_allclose(c[12], -r, atol=1e-12)
        np.testing.assert_allclose(c[16], fr[0] / 10.0, rtol=1e-12)  # contact.mu
        np.testing.assert_allclose(c[17:20], [fr[0], fr[0], fr[1]], rtol=1e-12)
        R0 = (1 - d) / d / mass
        np.testing.assert_allclose(R[4 * k], R0, rtol=1e-9)
        np.testing.assert_allclose(R[4 * k + 1: 4 * k + 3], R0 / 100, rtol=1e-9)
        np.testing.assert_allclose(R[4 * k + 3], R0 / 100 * fr[0] ** 2 / fr[1] ** 2, rtol=1e-9)
        np.testing.assert_allclose(aref[4 * k], K * d * r, rtol=1e-9)
        np.testing.assert_allclose(aref[4 * k + 1: 4 * k + 4], 0, atol=1e-12)
        np.testing.assert_allclose(pos[4 * k], -r, atol=1e-12)


def test_wall_cube_contacts_have_the_expected_geometry():
    """Cube face 1 mm inside the inner face of each rail: face contacts with the normal pointing from the rail (geom 1)
    into the pen towards the cube (geom 2), depth 1 mm, contact points on the overlap rectangle of the two faces
    (z between the floor-level bottom of the cube and the 12 mm top of the rail)."""
    cases = {22: ((-0.115 + 0.015 - 0.001, 0.135), (1, 0, 0)), 23: ((0.115 - 0.015 + 0.001, 0.135), (-1, 0, 0)),
             24: ((0.0, 0.10 + 0.015 - 0.001), (0, 1, 0)), 25: ((0.0, 0.17 - 0.015 + 0.001), (0, -1, 0))}
    for wall, (xy, normal) in cases.items():
        o = Oracle("push_loop", collision_mask=model.COLLIDE_WALL_CUBE)
        o.set_state(qpos=np.r_[np.zeros(6), xy, 0.015, 1, 0, 0, 0], qvel=np.zeros(12), ctrl=np.zeros(6))
        o.forward()
        con = o.get("contacts").reshape(-1, 27)
        assert len(con) == 4, (wall, len(con))
        for c in con:
            assert (int(c[14]), int(c[15]), int(c[13])) == (wall, 21, 4)
            np.testing.assert_allclose(c[3:6], normal, atol=1e-12)
            np.testing.assert_allclose(c[12], -0.001, atol=1e-12)
            assert -1e-12 <= c[2] <= 0.012 + 1e-12
            np.testing.assert_allclose(c[17:20], [1.5, 1.5, 1.5])  # the cube's friction (priority 1)
        # mid-surface points: half a depth inside the rail's face
        axis = int(np.argmax(np.abs(normal)))
        face = {22: -0.115, 23: 0.115, 24: 0.10, 25: 0.17}[wall]
        np.testing.assert_allclose(con[:, axis], face - 0.0005 * normal[axis], atol=1e-12)  # between the two overlapping faces


@pytest.mark.parametrize("task", ["push", "push_loop"])
def test_mpr_depth_is_bracketed_by_exact_hull_geometry(task):
    """Convex narrowphase (MPR) against plain numpy / Qhull on the hull vertices.  MPR reports the distance from the origin
    to the final portal triangle (a triangle inscribed in the boundary of the Minkowski difference A - B whose plane is
    within the tolerance of a supporting plane) and the direction to its closest point, so for every box-mesh / mesh-mesh
    contact: exact minimum penetration depth (nearest facet of the Minkowski difference) <= reported depth <= overlap of
    the two hulls along the reported normal, max_A a.n - min_B b.n; the normal is a unit vector from geom 1 to geom 2 and
    the contact point lies inside the overlap slab."""
    from scipy.spatial import ConvexHull

    rng = np.random.default_rng(31)
    m = model.load_compiled(task)
    nmesh = len(m["mesh_body"])
    lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
    hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)

    def world_verts(o, g, qpos):
        xpos, xmat = o.get("xpos").reshape(9, 3), o.get("xmat").reshape(9, 3, 3)
        if g < nmesh:
            b = int(m["mesh_body"][g])
            v = m["verts"][m["mesh_vertadr"][g]: m["mesh_vertadr"][g] + m["mesh_vertnum"][g]]
            return xpos[b] + v @ xmat[b].T
        c = g - nmesh - 1
        if c < int(m["ncube"]):
            return xpos[7 + c] + (corners * m["cube_size"][c]) @ xmat[7 + c].T
        w = c - int(m["ncube"])
        return m["wall_pos"][w] + corners * m["wall_size"][w]

    checked, exact_checked, tight = 0, 0, 0
    for trial in range(60):
        o = Oracle(task, collision_mask=model.COLLIDE_CUBE_MESH | model.COLLIDE_MESH_MESH | model.COLLIDE_WALL_MESH)
        qpos = np.zeros(13)
        qpos[:6] = rng.uniform(lo, hi)
        xp, _, _ = mjcf.arm_kinematics(m, qpos[:6])
        qpos[6:9] = xp[rng.integers(1, 7)] + rng.uniform(-0.03, 0.03, 3)
        q = rng.normal(size=4)
        qpos[9:13] = q / np.linalg.norm(q)
        o.set_state(qpos=qpos, qvel=np.zeros(12), ctrl=qpos[:6])
        o.forward()
        for c in o.get("contacts").reshape(-1, 27):
            g1, g2, n, depth = int(c[14]), int(c[15]), c[3:6], -c[12]
            A, B = world_verts(o, g1, qpos), world_verts(o, g2, qpos)
            assert abs(np.linalg.norm(n) - 1) < 1e-9 and depth > 0
            overlap = (A @ n).max() - (B @ n).min()
            assert depth <= overlap + 1e-6, (g1, g2, overlap, depth)
            tight += abs(overlap - depth) < 1e-5
            assert (B @ n).min() - 1e-6 <= c[0:3] @ n <= (A @ n).max() + 1e-6
            assert (A.mean(0) - B.mean(0)) @ n < overlap  # n points from geom 1 towards geom 2
            checked += 1
            if exact_checked < 12 and len(A) * len(B) < 60000:
                hull = ConvexHull((A[:, None, :] - B[None, :, :]).reshape(-1, 3))
                assert np.all(hull.equations[:, 3] < 1e-9)  # the origin is inside: the hulls do intersect
                assert -hull.equations[:, 3].max() <= depth + 1e-5, (g1, g2, -hull.equations[:, 3].max(), depth)
                exact_checked += 1
    assert checked >= 40 and exact_checked >= 8 and tight >= checked // 4, (checked, exact_checked, tight)


def test_free_cube_rotation_is_the_exact_quaternion_exponential():
    """Torque-free cube with isotropic inertia: the body-local angular velocity stays constant and the orientation after
    n substeps is exp(n h w / 2) exactly (each step multiplies by the exponential of h w, as mju_quatIntegrate does)."""
    o = Oracle("reach", collision_mask=0)
    w = np.array([1.0, -2.0, 3.0])
    qvel = np.zeros(12)
    qvel[9:12] = w
    o.set_state(qpos=np.r_[np.zeros(6), 0.0, 0.3, 0.5, 1, 0, 0, 0], qvel=qvel, ctrl=np.zeros(6))
    n = 100
    o.substep(n)
    st = o.get_state()
    ang = np.linalg.norm(w) * n * 0.002
    expect = np.r_[np.cos(ang / 2), np.sin(ang / 2) * w / np.linalg.norm(w)]
    np.testing.assert_allclose(st["qpos"][9:13], expect, atol=1e-13)
    np.testing.assert_allclose(st["qvel"][9:12], w, atol=1e-13)
    np.testing.assert_allclose(st["qvel"][6:9], [0, 0, -9.81 * n * 0.002], atol=1e-12)


def test_box_box_depth_is_the_exact_minimum_translation():
    """Cube-cube contacts (SAT over 15 axes + clipping) against Qhull: the deepest contact's depth is the exact minimum
    translation distance of the two boxes (nearest facet of their Minkowski difference), up to the 5 % preference the SAT
    gives face axes over edge axes, and the normal is a unit vector from cube 0 to cube 1 along which the boxes overlap
    by exactly that depth."""
    from scipy.spatial import ConvexHull

    rng = np.random.default_rng(13)
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float) * 0.015
    checked = 0
    for trial in range(60):
        o = Oracle("stack", collision_mask=model.COLLIDE_CUBE_CUBE)
        qpos = np.zeros(20)
        for c in range(2):
            q = rng.normal(size=4)
            qpos[9 + 7 * c: 13 + 7 * c] = q / np.linalg.norm(q)
        qpos[6:9] = [0.0, 0.2, 0.1]
        qpos[13:16] = qpos[6:9] + rng.uniform(-0.02, 0.02, 3)
        o.set_state(qpos=qpos, qvel=np.zeros(18), ctrl=np.zeros(6))
        o.forward()
        con = o.get("contacts").reshape(-1, 27)
        xpos, xmat = o.get("xpos").reshape(9, 3), o.get("xmat").reshape(9, 3, 3)
        A, B = xpos[7] + corners @ xmat[7].T, xpos[8] + corners @ xmat[8].T
        hull = ConvexHull((A[:, None, :] - B[None, :, :]).reshape(-1, 3))
        exact = -hull.equations[:, 3].max()  # > 0: distance from the origin (inside) to the nearest facet
        if exact <= 1e-9:
            assert len(con) == 0
            continue
        assert len(con) >= 1
        n = con[0, 3:6]
        overlap = (A @ n).max() - (B @ n).min()
        deepest = -con[:, 12].min()
        assert abs(np.linalg.norm(n) - 1) < 1e-12 and all(np.allclose(c[3:6], n) for c in con)
        assert exact - 1e-9 <= overlap <= 1.05 * exact + 1e-8, (exact, overlap)
        assert deepest <= overlap + 1e-9 and deepest >= 0.3 * overlap  # clipped points never deeper than the overlap
        checked += 1
    assert checked >= 40


def test_floor_contacts_are_hull_vertices_below_the_plane():
    """Plane-box and plane-hull contacts against plain numpy on the vertices: every floor contact sits at a vertex of the
    other geom that is below z = 0 (dist = its z, position = the vertex lifted by half the depth, normal +z), a geom has
    at most 4 of them, and its first one is the deepest vertex of a mesh / the contacts of a cube are all its corners
    below the plane."""
    rng = np.random.default_rng(19)
    m = model.load_compiled("push")
    nmesh = len(m["mesh_body"])
    lo = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
    hi = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float) * 0.015
    n_mesh_contacts = n_cube_contacts = 0
    for trial in range(40):
        o = Oracle("push", collision_mask=model.COLLIDE_FLOOR_CUBE | model.COLLIDE_FLOOR_MESH)
        qpos = np.zeros(13)
        qpos[:6] = rng.uniform(lo, hi)
        qpos[6:9] = [rng.uniform(-0.1, 0.1), rng.uniform(0.1, 0.25), rng.uniform(0.0, 0.02)]
        q = rng.normal(size=4)
        qpos[9:13] = q / np.linalg.norm(q)
        o.set_state(qpos=qpos, qvel=np.zeros(12), ctrl=qpos[:6])
        o.forward()
        if o.diag()["overflow"]:
            continue
        con = o.get("contacts").reshape(-1, 27)
        xpos, xmat = o.get("xpos").reshape(9, 3), o.get("xmat").reshape(9, 3, 3)
        for g in sorted(set(con[:, 15].astype(int))):
            cs = con[con[:, 15] == g]
            assert np.all(cs[:, 14] == nmesh) and len(cs) <= 4
            if g < nmesh:
                b = int(m["mesh_body"][g])
                V = xpos[b] + m["verts"][m["mesh_vertadr"][g]: m["mesh_vertadr"][g] + m["mesh_vertnum"][g]] @ xmat[b].T
                np.testing.assert_allclose(cs[0, 12], V[:, 2].min(), atol=1e-12)  # deepest vertex first
                n_mesh_contacts += len(cs)
            else:
                V = xpos[7] + corners @ xmat[7].T
                assert len(cs) == min(4, int((V[:, 2] < 0).sum()))
                n_cube_contacts += len(cs)
            for c in cs:
                np.testing.assert_allclose(c[3:6], [0, 0, 1], atol=1e-15)
                k = int(np.argmin(np.abs(V[:, 2] - c[12]) + np.linalg.norm(V[:, :2] - c[0:2], axis=1)))
                np.testing.assert_allclose(c[12], V[k, 2], atol=1e-12)
                np.testing.assert_allclose(c[0:3], V[k] - [0, 0, 0.5 * V[k, 2]], atol=1e-12)
                assert V[k, 2] < 0
    assert n_mesh_contacts >= 30 and n_cube_contacts >= 30, (n_mesh_contacts, n_cube_contacts)
