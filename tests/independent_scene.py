"""The scene of push_cube.xml as read by tools/make_independent_scene.py, with forward kinematics restated in numpy: shared by
tests/test_independent_pipeline.py (collision) and tests/test_independent_dynamics.py (dynamics, constraint rows, integration).
TEST INFRASTRUCTURE; shares no code or data with gym_lowcostrobot_b200/mjcf.py, the oracle or the kernels."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
S = np.load(os.path.join(HERE, "golden", "independent", "scene_push.npz"))
NB, NG = len(S["body_parent"]), len(S["geom_body"])
CORNERS = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
LO = np.array([-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533])
HI = np.array([3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599])


def quat_mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def axis_angle_mat(axis, angle):
    a = axis / np.linalg.norm(axis)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def body_frames(qpos):
    """(R[b], p[b]) of the 7 arm bodies and the world axes of the 6 hinge joints (anchored at the body origins)"""
    R, p, axes = [None] * NB, [None] * NB, []
    joint = 0
    for b in range(NB):
        par = S["body_parent"][b]
        Rp, pp = (np.eye(3), np.zeros(3)) if par < 0 else (R[par], p[par])
        Rb = quat_mat(S["body_quat"][b])
        if np.any(S["body_axis"][b]):
            axes.append(Rp @ Rb @ S["body_axis"][b])
            Rb = Rb @ axis_angle_mat(S["body_axis"][b], qpos[joint])
            joint += 1
        R[b], p[b] = Rp @ Rb, pp + Rp @ S["body_pos"][b]
    return R, p, axes


def world_geoms(qpos):
    """world-space vertex sets of the 20 mesh geoms and the cube (own forward kinematics on the raw mesh frames)"""
    R, p, _ = body_frames(qpos)
    verts = [p[S["geom_body"][g]] + S["hull_pts"][S["hull_adr"][g]:S["hull_adr"][g + 1]] @ R[S["geom_body"][g]].T for g in range(NG)]
    cube = qpos[6:9] + (CORNERS * S["cube_half"]) @ quat_mat(qpos[9:13]).T
    return verts, cube




def load_scene(task):
    """the fixture of another scene file (same arm; `ncube` free cubes, then static wall boxes)"""
    return np.load(os.path.join(HERE, "golden", "independent", f"scene_{task}.npz"))


def world_boxes(Sc, qpos):
    """world-space corner sets of the boxes of a scene: the free cubes (pose from qpos), then the static walls"""
    out = []
    ncube = int(Sc["ncube"])
    for c in range(ncube):
        p = qpos[6 + 7 * c: 13 + 7 * c]
        out.append(p[:3] + (CORNERS * Sc["box_half"][c]) @ quat_mat(p[3:7]).T)
    for w, pos in enumerate(Sc["wall_pos"]):
        out.append(pos + CORNERS * Sc["box_half"][ncube + w])
    return out
