"""A second, structurally different restatement of ONE mj_step of the PushCube scene in numpy, against the oracle.

Everything the step needs is rebuilt from tests/golden/independent/scene_push.npz -- the raw MJCF numbers read by
tools/make_independent_scene.py's own reader (masses, inertial frames, armature, damping, force ranges, kp / kv, per-geom
friction / condim / priority / solref / solimp with the default classes resolved there) -- and from the published formulation
of MuJoCo's computation pipeline, NOT from gym_lowcostrobot_b200/mjcf.py (the compiler the oracle and the kernels share) nor
from the oracle's intermediate arrays:

  mass matrix          sum over bodies of m Jv^T Jv + Jw^T I Jw at the body's centre of mass, + armature
  bias forces          Lagrange: d/dt(M) v - 1/2 grad_q(v^T M v) + grad_q V, by central differences of the own M(q), V(q)
  smooth forces        position servo kp (clamp(ctrl, joint range: inheritrange=1) - q) - kv v clamped to the joint's
                       actuatorfrcrange, joint damping
  model constants      dof_invweight0 / body_invweight0 from the own M^-1 at qpos0
  constraint rows      joint limits and the contacts of the oracle's list (positions / frames / distances / geom ids only: the
                       geometry is checked by test_independent_pipeline.py): contact parameter mixing from the raw geom
                       attributes, Jacobians from the own kinematics, impedance d(r), R, D, K, B, aref, the elliptic cone's
                       regularised mu and R scaling
  solve                the oracle's qacc must be a stationary point and a minimiser of the OWN primal objective
  integration          implicitfast velocity update (M - h dF/dv) and position / quaternion integration

so a wrong number out of the shared model compiler, a wrong Jacobian, row parameter or solver result in the oracle shows up
as a disagreement here.  States are harvested from oracle rollouts with random actions (plus constructed contact-rich ones).
Reference: `mujoco.mj_step` (reach_cube_env.py:276-277); SURVEY Appendix A.3.
"""
import numpy as np
import pytest
from scipy.optimize import minimize

from independent_scene import HI, LO, NB, NG, body_frames, load_scene, quat_mat
from oracle.oracle import Oracle

S = load_scene("push")
NCUBE, NV, NQ = 1, 12, 13
H, G = float(S["timestep"]), 9.81
IMPRATIO = float(S["impratio"])
MINVAL, MINIMP, MAXIMP, MINMU = 1e-15, 1e-4, 0.9999, 1e-5
ARM = S["joints"]  # armature, damping, frcrange lo hi, range lo hi
DOF_INVW = BODY_INVW = None


def use_scene(task):
    """switch the module to another scene fixture (same arm; bodies 7.. = the free cubes)"""
    global S, NCUBE, NV, NQ, H, IMPRATIO, ARM, DOF_INVW, BODY_INVW
    S = load_scene(task)
    NCUBE = int(S["ncube"])
    NV, NQ = 6 + 6 * NCUBE, 6 + 7 * NCUBE
    H, IMPRATIO, ARM = float(S["timestep"]), float(S["impratio"]), S["joints"]
    DOF_INVW, BODY_INVW = invweights()


# ------------------------------------------------------------------------------------------------ kinematics, M, bias
def point_jacobian(qpos, body, point):
    """(Jp, Jr) 3 x NV of a world point attached to `body` (1..6 arm links, 0 = welded base, 7.. = the cubes, -1 = world)"""
    Jp, Jr = np.zeros((3, NV)), np.zeros((3, NV))
    if body >= 7:
        c = body - 7
        R = quat_mat(qpos[9 + 7 * c:13 + 7 * c])
        d0 = 6 + 6 * c
        Jp[:, d0:d0 + 3] = np.eye(3)
        for k in range(3):  # rotational dofs of a free joint are in the body frame
            Jp[:, d0 + 3 + k] = np.cross(R[:, k], point - qpos[6 + 7 * c:9 + 7 * c])
            Jr[:, d0 + 3 + k] = R[:, k]
    elif body >= 1:
        _, p, axes = body_frames(qpos)
        for j in range(body):  # joint j moves bodies j + 1 .. 6 (a serial chain)
            Jp[:, j] = np.cross(axes[j], point - p[j + 1])
            Jr[:, j] = axes[j]
    return Jp, Jr


def mass_matrix(qpos):
    R, p, _ = body_frames(qpos)
    M = np.diag(np.r_[ARM[:, 0], np.concatenate([np.r_[np.full(3, S["cube_masses"][c]), S["cube_diaginertias"][c]] for c in range(NCUBE)])])
    for b in range(1, NB):
        ine = S["inertial"][b]
        com = p[b] + R[b] @ ine[0:3]
        Ri = R[b] @ quat_mat(ine[3:7])
        Iw = Ri @ np.diag(ine[8:11]) @ Ri.T
        Jp, Jr = point_jacobian(qpos, b, com)
        M += ine[7] * Jp.T @ Jp + Jr.T @ Iw @ Jr
    return M


def potential(qpos):
    R, p, _ = body_frames(qpos)
    return sum(S["inertial"][b][7] * G * (p[b] + R[b] @ S["inertial"][b][0:3])[2] for b in range(1, NB))


def bias_forces(qpos, qvel, eps=1e-6):
    """Coriolis / centrifugal / gravity of the arm by finite differences of the own M(q), V(q); the cube's inertia is isotropic,
    so its bias is its weight"""
    c = np.zeros(NV)
    v = qvel[:6]
    dM = []
    for k in range(6):
        e = np.zeros(NQ)
        e[k] = eps
        dM.append((mass_matrix(qpos + e)[:6, :6] - mass_matrix(qpos - e)[:6, :6]) / (2 * eps))
        c[k] = (potential(qpos + e) - potential(qpos - e)) / (2 * eps) - 0.5 * v @ dM[k] @ v
    c[:6] += sum(dM[k] * v[k] for k in range(6)) @ v
    for cb in range(NCUBE):
        c[6 + 6 * cb + 2] = float(S["cube_masses"][cb]) * G
    return c


def smooth_acceleration(qpos, qvel, ctrl, M):
    # <position ... inheritrange="1"/>: the actuator's ctrlrange is the joint's range, and ctrl is clamped to it
    u = np.clip(ctrl, ARM[:, 4], ARM[:, 5])
    act = np.clip(float(S["act_kp"]) * (u - qpos[:6]) - float(S["act_kv"]) * qvel[:6], ARM[:, 2], ARM[:, 3])
    f = -bias_forces(qpos, qvel)
    f[:6] += act - ARM[:, 1] * qvel[:6]
    return np.linalg.solve(M, f)


def invweights():
    """dof_invweight0 of the 6 hinges and body_invweight0 (translation, rotation) of bodies 0..6 and the cubes at qpos0"""
    q0 = np.r_[np.zeros(6), np.concatenate([np.r_[S["cube_pos0s"][c], 1, 0, 0, 0] for c in range(NCUBE)])]
    Minv = np.linalg.inv(mass_matrix(q0))
    R, p, _ = body_frames(q0)
    body = np.zeros((7 + NCUBE, 2))
    for b in range(1, 7 + NCUBE):
        com = q0[6 + 7 * (b - 7):9 + 7 * (b - 7)] if b >= 7 else p[b] + R[b] @ S["inertial"][b][0:3]
        Jp, Jr = point_jacobian(q0, b, com)
        body[b] = np.trace(Jp @ Minv @ Jp.T) / 3, np.trace(Jr @ Minv @ Jr.T) / 3
    return np.diag(Minv)[:6].copy(), body


use_scene("push")


# ------------------------------------------------------------------------------------------------ constraint rows
def mix(g1, g2):
    """MuJoCo's contact parameter mixing of two geoms: friction5, condim, solref, solimp"""
    a, b = S["geom_par"][g1], S["geom_par"][g2]
    if a[4] != b[4]:
        w = a if a[4] > b[4] else b
        fr, dim, solref, solimp = w[0:3], int(w[3]), w[5:7], w[7:12]
    else:
        mixw = a[12] / (a[12] + b[12])
        fr, dim = np.maximum(a[0:3], b[0:3]), int(max(a[3], b[3]))
        # solref: weighted like solimp when both are in the (timeconst, dampratio) form, otherwise the element-wise minimum
        solref = mixw * a[5:7] + (1 - mixw) * b[5:7] if a[5] > 0 and b[5] > 0 else np.minimum(a[5:7], b[5:7])
        solimp = mixw * a[7:12] + (1 - mixw) * b[7:12]
    fr = np.maximum(MINMU, fr)
    return np.array([fr[0], fr[0], fr[1], fr[2], fr[2]]), dim, solref, solimp


def impedance(solimp, dist):
    dmin, dmax = np.clip(solimp[0:2], MINIMP, MAXIMP)
    width, mid, power = max(0.0, solimp[2]), np.clip(solimp[3], MINIMP, MAXIMP), max(1.0, solimp[4])
    if dmin == dmax or width <= MINVAL:
        return 0.5 * (dmin + dmax), dmax
    x = abs(dist) / width
    if x >= 1:
        return dmax, dmax
    if x == 0:
        return dmin, dmax
    y = x ** power / mid ** (power - 1) if x <= mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
    return dmin + y * (dmax - dmin), dmax


def stiffness_damping(solref, dmax):
    if solref[0] <= 0:  # direct form (-stiffness, -damping); PushCubeLoop's floor has solref = "0 0": no reference acceleration at all
        return -solref[0] / max(MINVAL, dmax * dmax), -solref[1] / max(MINVAL, dmax)
    tc, dr = max(solref[0], 2 * H), solref[1]  # refsafe
    return 1 / max(MINVAL, dmax * dmax * tc * tc * dr * dr), 2 / max(MINVAL, dmax * tc)


def geom_body(g):
    if g < NG:
        return int(S["geom_body"][g])
    return 7 + (g - NG - 1) if NG < g <= NG + NCUBE else -1  # the floor and the static walls belong to the world


def constraint_rows(qpos, qvel, contacts):
    """J, aref, D of the joint-limit rows followed by the rows of `contacts` (oracle records: pos, frame, dist, geom ids),
    and per contact (dim, regularised mu, friction5) for the objective"""
    J, aref, D, units = [], [], [], []
    for j in range(6):
        for side, dist in ((1.0, qpos[j] - ARM[j, 4]), (-1.0, ARM[j, 5] - qpos[j])):
            if dist < 0:
                row = np.zeros(NV)
                row[j] = side
                imp, dmax = impedance(np.array([0.9, 0.95, 0.001, 0.5, 2]), dist)
                K, B = stiffness_damping(np.array([0.02, 1.0]), dmax)
                J.append(row)
                D.append(1 / max(MINVAL, (1 - imp) / imp * DOF_INVW[j]))
                aref.append(-B * (row @ qvel) - K * imp * dist)
    nlim = len(J)
    for c in contacts:
        pos, frame, dist, g1, g2 = c[0:3], c[3:12].reshape(3, 3), c[12], int(c[14]), int(c[15])
        fr, dim, solref, solimp = mix(g1, g2)
        b1, b2 = geom_body(g1), geom_body(g2)
        Jp1, Jr1 = point_jacobian(qpos, b1, pos)
        Jp2, Jr2 = point_jacobian(qpos, b2, pos)
        rows = [frame[k] @ (Jp2 - Jp1) for k in range(3)] + [frame[k] @ (Jr2 - Jr1) for k in range(3)]
        rows = rows[:dim]
        imp, dmax = impedance(solimp, dist)
        K, B = stiffness_damping(solref, dmax)
        tran = BODY_INVW[max(b1, 0), 0] * (b1 > 0) + BODY_INVW[max(b2, 0), 0] * (b2 > 0)
        R = np.zeros(dim)
        R[0] = max(MINVAL, (1 - imp) / imp * tran)
        # elliptic cone: the friction rows are regularised relative to the normal row, R1 = R0 / impratio, and scaled so that
        # R_j mu_j^2 is the same for all friction dimensions; the cone's mu follows from R1 / R0
        R[1] = R[0] / max(MINVAL, IMPRATIO)
        for k in range(2, dim):
            R[k] = R[1] * fr[0] ** 2 / fr[k - 1] ** 2
        mu = fr[0] * np.sqrt(R[1] / R[0])
        for k in range(dim):
            J.append(rows[k])
            D.append(1 / R[k])
            aref.append(-B * (rows[k] @ qvel) - (K * imp * dist if k == 0 else 0.0))
        units.append((dim, mu, fr))
    return np.array(J).reshape(-1, NV), np.array(aref), np.array(D), nlim, units


def objective(a, M, a0, J, aref, D, nlim, units):
    """MuJoCo's primal objective with elliptic cones: Gauss term + sum of the row / cone costs, and its gradient"""
    jar = J @ a - aref
    da = a - a0
    cost, g = 0.5 * da @ M @ da, np.zeros(len(jar))
    for i in range(nlim):
        if jar[i] < 0:
            cost += 0.5 * D[i] * jar[i] ** 2
            g[i] = D[i] * jar[i]
    i = nlim
    for dim, mu, fr in units:
        x = jar[i:i + dim]
        u = x * np.r_[mu, fr[:dim - 1]]
        N, T = u[0], np.linalg.norm(u[1:])
        if N >= mu * T or (T <= 0 and N >= 0):
            pass  # inside the dual cone: no force
        elif mu * N + T <= 0 or (T <= 0 and N < 0):
            cost += 0.5 * np.sum(D[i:i + dim] * x * x)
            g[i:i + dim] = D[i:i + dim] * x
        else:
            dm = D[i] / (mu * mu * (1 + mu * mu))
            nmt = N - mu * T
            cost += 0.5 * dm * nmt ** 2
            g[i] = dm * nmt * mu
            g[i + 1:i + dim] = -dm * nmt * mu / T * u[1:] * fr[:dim - 1]
        i += dim
    return cost, M @ da + J.T @ g


def integrate(qpos, qvel, qacc, M):
    """implicitfast: (M - h dF/dv) dv = h M qacc with dF/dv = -(damping + kv) on the arm dofs; then positions"""
    Dv = np.zeros(NV)
    Dv[:6] = -(ARM[:, 1] + float(S["act_kv"]))
    v = qvel + H * np.linalg.solve(M - H * np.diag(Dv), M @ qacc)
    q = qpos.copy()
    q[:6] += H * v[:6]
    for c in range(NCUBE):
        p0, d0 = 6 + 7 * c, 6 + 6 * c
        q[p0:p0 + 3] += H * v[d0:d0 + 3]
        w = v[d0 + 3:d0 + 6]
        ang = np.linalg.norm(w) * H
        if ang > 0:
            ax = w / np.linalg.norm(w)
            a0, b0 = qpos[p0 + 3:p0 + 7], np.r_[np.cos(ang / 2), np.sin(ang / 2) * ax]
            q[p0 + 3:p0 + 7] = np.r_[a0[0] * b0[0] - a0[1:] @ b0[1:], a0[0] * b0[1:] + b0[0] * a0[1:] + np.cross(a0[1:], b0[1:])]
            q[p0 + 3:p0 + 7] /= np.linalg.norm(q[p0 + 3:p0 + 7])
    return q, v


# ------------------------------------------------------------------------------------------------ states
def harvested_states(task):
    """mid-episode states of oracle rollouts with random actions, and constructed contact-rich ones"""
    rng = np.random.default_rng(17)
    out = []
    for seed in range(4):
        o = Oracle(task)
        o.reset(seed=seed)
        for t in range(18):
            o.step(rng.uniform(-1, 1, o.na).astype(np.float32))
            if t in (5, 11, 17):
                st = o.get_state()
                out.append((st["qpos"], st["qvel"], st["ctrl"]))
    for k in range(10):
        q = np.zeros(NQ)
        q[:6] = rng.uniform(LO, HI)
        if k % 2:
            q[1], q[2] = rng.uniform(0.8, 1.22), rng.uniform(1.0, 1.74)  # arm in the floor
        if k == 4:
            q[5] = 0.05  # beyond the range of joint_6: a limit row
        R, p, _ = body_frames(q)
        q[6:9] = p[rng.integers(4, 7)] + rng.uniform(-0.03, 0.03, 3) if k % 3 else [rng.uniform(-0.1, 0.1), rng.uniform(0.1, 0.25), 0.013]
        quat = rng.normal(size=4)
        q[9:13] = quat / np.linalg.norm(quat)
        if task == "push_loop" and k % 2 == 0:  # the cube against / inside one of the rails
            wl = rng.integers(0, 4)
            q[6:9] = S["wall_pos"][wl] + rng.uniform(-1, 1, 3) * (S["box_half"][1 + wl] + 0.01) + [0, 0, 0.008]
        if NCUBE == 2:  # the second cube on / inside / beside the first (cube-cube contacts)
            q[13:16] = q[6:9] + rng.uniform(-0.02, 0.02, 3) + [0, 0, 0.02 * (k % 2)]
            quat = rng.normal(size=4)
            q[16:20] = quat / np.linalg.norm(quat) if k % 3 else q[9:13]
        out.append((q, rng.normal(scale=0.5, size=NV), rng.uniform(LO, HI)))
    return out


@pytest.mark.parametrize("task", ["push", "stack", "push_loop"])
def test_one_substep_from_an_independent_restatement(task):
    use_scene(task)
    checked = rows_checked = cone_states = limit_rows = cube_cube = wall_rows = 0
    for qpos, qvel, ctrl in harvested_states(task):
        o = Oracle(task)
        o.set_state(qpos=qpos, qvel=qvel, ctrl=ctrl, warm=np.zeros(NV))
        o.forward()
        d = o.diag()
        assert d["overflow"] == 0
        # (1) inertia and smooth dynamics
        M = mass_matrix(qpos)
        np.testing.assert_allclose(o.get("M").reshape(NV, NV), M, rtol=0, atol=1e-10)
        a0 = smooth_acceleration(qpos, qvel, ctrl, M)
        np.testing.assert_allclose(o.get("qacc_smooth"), a0, rtol=1e-6, atol=1e-5)
        # (2) constraint rows from the raw MJCF numbers
        con = o.get("contacts").reshape(-1, 27)
        J, aref, D, nlim, units = constraint_rows(qpos, qvel, con)
        nefc = d["nefc"]
        assert len(aref) == nefc, (len(aref), nefc)
        if nefc:
            np.testing.assert_allclose(o.get("efc_J").reshape(nefc, NV), J, rtol=0, atol=1e-9)
            np.testing.assert_allclose(o.get("efc_D"), D, rtol=1e-9)
            np.testing.assert_allclose(o.get("efc_aref"), aref, rtol=1e-8, atol=1e-7)
            for c, (dim, mu, fr) in zip(con, units):
                assert int(c[13]) == dim and abs(c[16] - mu) < 1e-12 * max(1, mu) and np.allclose(c[17:22], fr, rtol=1e-12)
                cube_cube += min(int(c[14]), int(c[15])) > NG
                wall_rows += max(int(c[14]), int(c[15])) > NG + NCUBE
            rows_checked += nefc
            limit_rows += nlim
        # (3) the oracle's qacc is a stationary point and a minimiser of the own objective (own a0: the oracle's agrees to 1e-6)
        a = o.get("qacc")
        if nefc:
            a0o = o.get("qacc_smooth")
            f = lambda x: objective(x, M, a0o, J, aref, D, nlim, units)
            cost, grad = f(a)
            scale = np.linalg.norm(M @ (a - a0o)) + 1e-3
            assert np.linalg.norm(grad) <= 1e-5 * scale, (np.linalg.norm(grad), scale, d)
            best = minimize(lambda x: f(x)[0], a, jac=lambda x: f(x)[1], method="BFGS", options={"gtol": 1e-12, "maxiter": 500}).fun
            assert cost <= best + 1e-7 * (abs(cost) + 1e-3)
            cone_states += cost > 1e-9
        else:
            np.testing.assert_allclose(a, a0, rtol=1e-6, atol=1e-5)
        # (4) integration
        q1, v1 = integrate(qpos, qvel, a, M)
        o.substep(1)
        st = o.get_state()
        np.testing.assert_allclose(st["qvel"], v1, rtol=0, atol=1e-9 * max(1.0, np.abs(v1).max()))
        np.testing.assert_allclose(st["qpos"], q1, rtol=0, atol=1e-11)
        checked += 1
    assert checked >= 20 and rows_checked >= 300 and cone_states >= 10 and limit_rows >= 1, (checked, rows_checked, cone_states, limit_rows)
    assert task != "stack" or cube_cube >= 8, cube_cube
    assert task != "push_loop" or wall_rows >= 8, wall_rows
