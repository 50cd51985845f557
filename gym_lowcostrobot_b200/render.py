"""Image observations: the host side of ``lcr_body_poses`` + ``lcr_render`` (csrc/lcr_render.cu).

Replaces ``mujoco.Renderer(model)`` + ``update_scene(data, camera=...)`` + ``render()`` of the reference
(``reach_cube_env.py:109-112,288-292`` and the same lines of the other envs): two fixed cameras (``camera_front``,
``camera_top``, identical in all six scene files), 240 x 320 RGB uint8, for the whole batch in two launches.  The scene is the
floor, the convex hulls of the arm's VISUAL meshes (MuJoCo draws geom groups 0-2: the ``visual`` class, not the ``collision``
one) with the materials of ``follower.xml`` (white 0.8 links, black 0.1 motors), and the boxes (red / blue cubes, white rails).
See the kernel's header for what is not reproduced of the OpenGL renderer.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi

HEIGHT, WIDTH = 240, 320  # reach_cube_env.py:110-111
GEOM_WORDS = 16
FOVY_DEG = 45.0  # MuJoCo's default camera fovy; the scene files do not set one


def _camera(pos, x_axis, y_axis):
    x = np.asarray(x_axis, np.float64)
    x /= np.linalg.norm(x)
    y = np.asarray(y_axis, np.float64)
    y -= x * (x @ y)  # MuJoCo orthogonalises xyaxes the same way
    y /= np.linalg.norm(y)
    R = np.stack([x, y, np.cross(x, y)], axis=1)  # columns: camera x, y, z (the camera looks along -z)
    return np.r_[np.asarray(pos, np.float64), R.reshape(-1), FOVY_DEG]


# <camera name="camera_front" pos="0.049 0.5 0.225" xyaxes="-0.998 0.056 -0.000 -0.019 -0.335 0.942"/>
# <camera name="camera_top" pos="0 0.1 0.6" euler="0 0 0"/>   (identity orientation: looks straight down, image up = world +y)
CAMERAS = {"camera_front": _camera([0.049, 0.5, 0.225], [-0.998, 0.056, -0.000], [-0.019, -0.335, 0.942]),
           "camera_top": _camera([0.0, 0.1, 0.6], [1, 0, 0], [0, 1, 0])}
CUBE_RGB = {"stack": ([0.5, 0, 0], [0, 0, 0.5])}  # stack_two_cubes.xml: cube_red, cube_blue; every other scene: one red cube


def hull_planes(verts):
    """unique half-spaces n . x + d <= 0 of the convex hull of ``verts`` (Qhull), largest faces first: a ray that misses the hull
    leaves the clipping loop as soon as an entry crossing lies behind an exit crossing, which the big faces decide soonest"""
    from scipy.spatial import ConvexHull

    v = np.asarray(verts, np.float64)
    hull = ConvexHull(v)
    tri = v[hull.simplices]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    key, inv = np.unique(np.round(hull.equations, 7), axis=0, return_inverse=True)
    total = np.bincount(inv.reshape(-1), weights=area, minlength=len(key))
    first = np.array([np.flatnonzero(inv.reshape(-1) == k)[0] for k in range(len(key))])
    order = np.argsort(-total, kind="stable")
    return hull.equations[first[order]]


def build_scene(compiled, task):
    """(geoms [G, 16] float32, planes [P, 4] float32) for ``lcr_render`` from the compiled model constants"""
    m = compiled
    names = [str(x) for x in m["mesh_names"]]
    geoms, planes = [], []
    for g, name in enumerate(names):
        if name.endswith("_collision"):
            continue  # geom group 3: not drawn
        v = m["verts"][int(m["mesh_vertadr"][g]): int(m["mesh_vertadr"][g]) + int(m["mesh_vertnum"][g])]
        pl = hull_planes(v)
        rgb = [0.1] * 3 if name.endswith("_motor") else [0.8] * 3
        # (mesh_center / mesh_half: the hull's bounding box in the body frame, mesh_rbound its bounding sphere about the same centre)
        geoms.append([0, int(m["mesh_body"][g]), sum(len(p) for p in planes), len(pl), *m["mesh_center"][g], float(m["mesh_rbound"][g]),
                      *m["mesh_half"][g], *rgb, 0, 0])
        planes.append(pl)
    ncube = int(m["ncube"])
    colors = CUBE_RGB.get(task, ([0.5, 0, 0],))
    for c in range(ncube):
        h = np.asarray(m["cube_size"][c], np.float64)
        geoms.append([1, 7 + c, 0, 0, 0, 0, 0, float(np.linalg.norm(h)), *h, *colors[c], 0, 0])
    if "wall_size" in m:
        for w, h in enumerate(np.asarray(m["wall_size"], np.float64)):
            geoms.append([1, 7 + ncube + w, 0, 0, 0, 0, 0, float(np.linalg.norm(h)), *h, 1, 1, 1, 0, 0])
    return np.asarray(geoms, np.float32), np.concatenate(planes).astype(np.float32)


class BatchRenderer:
    """Renders ``image_front`` / ``image_top`` ``[num_envs, 240, 320, 3]`` uint8 device tensors of a batched env."""

    def __init__(self, env, height=HEIGHT, width=WIDTH, cameras=("camera_front", "camera_top")):
        self.env, self.h, self.w, self.names = env, int(height), int(width), tuple(cameras)
        geoms, planes = build_scene(env.compiled, env.task)
        dev = env.device
        self.geoms, self.planes = torch.from_numpy(geoms).to(dev), torch.from_numpy(planes).to(dev)
        self.cams = np.ascontiguousarray(np.stack([CAMERAS[c] for c in self.names]), np.float32)
        self.nslot = int(env._L.lcr_pose_slots(env._h))
        self.poses = torch.empty(env.num_envs, self.nslot, 12, dtype=torch.float32, device=dev)
        self.images = torch.empty(env.num_envs, len(self.names), self.h, self.w, 3, dtype=torch.uint8, device=dev)

    def render(self):
        """{"image_front": [n, H, W, 3] uint8, ...}: views of one internal buffer, overwritten by the next call"""
        env, p = self.env, (lambda t: C.c_void_p(t.data_ptr()))
        with torch.cuda.device(env.device):
            st = env._stream()
            capi.check(env._L.lcr_body_poses(env._h, p(self.poses), st))
            if env._L.lcr_render(p(self.poses), env.num_envs, self.nslot, p(self.geoms), self.geoms.shape[0], p(self.planes),
                                 self.cams.ctypes.data_as(C.c_void_p), len(self.names), self.h, self.w, p(self.images), st):
                raise capi.LcrError("lcr_render failed (invalid arguments or launch error)")
        return {"image_" + n.split("_", 1)[1]: self.images[:, k] for k, n in enumerate(self.names)}
