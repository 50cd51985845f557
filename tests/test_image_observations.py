"""Image observations (reach_cube_env.py:109-112,288-292; the reference's own test runs observation_mode="image" too,
tests/test_env.py:9-12): API shape of the three observation modes, and the CUDA ray caster (csrc/lcr_render.cu) against an
independent numpy restatement of the same camera model and convex clipping (float64, vectorised over the pixels).

MuJoCo's OpenGL renderer is not available here and could not be matched pixel for pixel anyway (original concave visual
meshes, shadows, anti-aliasing); what is pinned is the geometry: camera poses and field of view from the scene files, body
poses from the simulator state, hull silhouettes, box faces, the floor checker, the two-light shading.  Tolerance: the two
images must agree within 2 grey levels on at least 99 % of the pixels (silhouette-edge pixels may fall on the other side of
a float32 rounding) and the mean absolute difference must be below 0.5 grey levels.
"""
import numpy as np
import pytest

from gym_lowcostrobot_b200 import model, render


def test_scene_tables_and_cameras():
    for task, nbox in (("push", 1), ("stack", 2), ("push_loop", 5)):
        geoms, planes = render.build_scene(model.load_compiled(task), task)
        assert geoms.shape == (13 + nbox, render.GEOM_WORDS) and planes.shape[1] == 4  # 13 visual meshes (7 links + 6 motors)
        hulls = geoms[geoms[:, 0] == 0]
        assert int(hulls[:, 3].sum()) == len(planes) and np.allclose(np.linalg.norm(planes[:, :3], axis=1), 1, atol=1e-5)
        assert sorted(float(x) for x in np.unique(hulls[:, 11].round(3))) == pytest.approx([0.1, 0.8])  # black motors, white links
        # every hull's bounding-sphere centre is inside all of its half-spaces
        for g in hulls:
            pl = planes[int(g[2]): int(g[2]) + int(g[3])]
            assert np.all(pl[:, :3] @ g[4:7] + pl[:, 3] < 1e-6)
    front, top = render.CAMERAS["camera_front"], render.CAMERAS["camera_top"]
    R = front[3:12].reshape(3, 3)
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1)
    assert -R[:, 2] @ (np.array([0, 0.15, 0.0]) - front[:3]) > 0  # looks towards the workspace in front of the arm
    assert np.allclose(top[3:12].reshape(3, 3), np.eye(3)) and top[2] == 0.6  # straight down from 0.6 m


def reference_image(poses, geoms, planes, cam, H, W):
    """numpy restatement of the ray caster for ONE env and camera: float64, all pixels at once"""
    pos, R, fovy = cam[:3], cam[3:12].reshape(3, 3), cam[12]
    th = np.tan(np.radians(fovy) / 2)
    jj, ii = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    dc = np.stack([(2 * jj / W - 1) * th * W / H, (1 - 2 * ii / H) * th, -np.ones_like(jj)], -1)
    d = dc @ R.T
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    tbest = np.where(d[..., 2] < -1e-9, -pos[2] / np.minimum(d[..., 2], -1e-30), np.inf)
    hit = np.where(np.isfinite(tbest), -2, -1)
    nrm = np.zeros_like(d)
    nrm[..., 2] = 1
    for gi, g in enumerate(geoms):
        P = poses[int(g[1])]
        p, Rb = P[:3], P[3:].reshape(3, 3)
        ol, dl = Rb.T @ (pos - p), d @ Rb
        if g[0] == 1:
            pl = np.concatenate([np.c_[np.eye(3), -g[8:11]], np.c_[-np.eye(3), -g[8:11]]])
        else:
            pl = planes[int(g[2]): int(g[2]) + int(g[3])]
        den = dl @ pl[:, :3].T                       # [H, W, P]
        dist = pl[:, :3] @ ol + pl[:, 3]             # [P]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = -dist / den
        t_in = np.where(den < -1e-12, t, -np.inf)
        t_out = np.where(den > 1e-12, t, np.inf)
        t0, k0 = t_in.max(-1), t_in.argmax(-1)
        t1 = t_out.min(-1)
        outside_parallel = ((np.abs(den) <= 1e-12) & (dist > 0)).any(-1)
        ok = (t0 <= np.minimum(t1, tbest)) & (t0 > 0) & (t0 < tbest) & ~outside_parallel
        tbest = np.where(ok, t0, tbest)
        hit = np.where(ok, gi, hit)
        nrm = np.where(ok[..., None], pl[k0, :3] @ Rb.T, nrm)
    hp = pos + tbest[..., None] * d
    with np.errstate(invalid="ignore"):
        odd = ((np.floor(hp[..., 0] * 10) + np.floor(hp[..., 1] * 10)) % 2) != 0
    base = np.where(odd[..., None], [0.1, 0.2, 0.3], [0.2, 0.3, 0.4])
    for gi, g in enumerate(geoms):
        base = np.where((hit == gi)[..., None], g[11:14], base)
    l2 = np.array([0, 0, 3.0]) - hp
    with np.errstate(invalid="ignore"):
        l2 /= np.linalg.norm(l2, axis=-1, keepdims=True)
        lum = 0.3 + 0.6 * np.maximum(0, -(nrm * d).sum(-1)) + 0.7 * np.maximum(0, (nrm * l2).sum(-1))
    rgb = np.minimum(1, base * lum[..., None])
    sky = np.maximum(0, d[..., 2])[..., None] * np.array([0.3, 0.5, 0.7])
    rgb = np.where((hit == -1)[..., None], sky, rgb)
    return np.floor(rgb * 255 + 0.5).astype(np.uint8), hit


def test_numpy_ray_caster_sees_the_scene():
    """the restatement itself, on CPU: with the arm at rest the front camera sees floor, white link hulls, black motors and the
    red cube where the camera model says it is"""
    from gym_lowcostrobot_b200 import mjcf

    m = model.load_compiled("push")
    geoms, planes = render.build_scene(m, "push")
    xpos, xmat, _ = mjcf.arm_kinematics(m, np.zeros(6))
    cube = np.array([0.12, 0.14, 0.015])  # beside the arm, not under it
    poses = np.zeros((8, 12))
    poses[:7, :3], poses[:7, 3:] = xpos[:7], np.asarray(xmat)[:7].reshape(7, 9)
    poses[7, :3], poses[7, 3:] = cube, np.eye(3).reshape(-1)
    H, W = 120, 160
    for name in ("camera_front", "camera_top"):
        cam = render.CAMERAS[name]
        img, hit = reference_image(poses, geoms.astype(np.float64), planes.astype(np.float64), cam, H, W)
        assert (hit == -2).mean() > 0.4 and (hit >= 0).sum() > 200       # mostly floor, the arm covers a few hundred pixels
        red = hit == 13
        assert red.sum() >= 6 and img[red][:, 0].min() > img[red][:, 1].max()   # the cube is red
        R = cam[3:12].reshape(3, 3)
        c = R.T @ (cube - cam[:3])                                       # cube centre in the camera frame -> expected pixel
        th = np.tan(np.radians(cam[12]) / 2)
        u, v = (c[0] / -c[2] / (th * W / H) + 1) * W / 2, (1 - c[1] / -c[2] / th) * H / 2
        ys, xs = np.nonzero(red)
        assert abs(xs.mean() + 0.5 - u) < 2.0 and abs(ys.mean() + 0.5 - v) < 2.0, (name, xs.mean(), u, ys.mean(), v)
        whites, blacks = (hit >= 0) & (hit < 13) & (img[..., 0] > 120), (hit >= 0) & (hit < 13) & (img[..., 0] < 60)
        assert whites.sum() > 50 and blacks.sum() > 20


@pytest.mark.gpu
def test_observation_modes_and_images_against_the_numpy_ray_caster():
    import torch

    import gym_lowcostrobot_b200 as glr

    n = 3
    for env_id, keys_image, keys_both in (
            ("PushCube-v0", ["arm_qpos", "arm_qvel", "target_pos", "image_front", "image_top"],
             ["arm_qpos", "arm_qvel", "target_pos", "image_front", "image_top", "cube_pos"]),
            ("StackTwoCubes-v0", ["arm_qpos", "arm_qvel", "image_front", "image_top"],
             ["arm_qpos", "arm_qvel", "image_front", "image_top", "cube_red_pos", "cube_blue_pos"])):
        env = glr.make(env_id, num_envs=n, observation_mode="image")
        obs, _ = env.reset(seed=5)
        assert list(obs) == keys_image and list(env.single_observation_space.spaces) == keys_image
        assert obs["image_front"].shape == (n, 240, 320, 3) and obs["image_front"].dtype == torch.uint8 and obs["image_top"].is_cuda
        env.close()
        env = glr.make(env_id, num_envs=n, observation_mode="both")
        obs, _ = env.reset(seed=5)
        assert list(obs) == keys_both
        g = torch.Generator(device="cuda").manual_seed(1)
        for _ in range(4):
            obs, r, te, tr, info = env.step(torch.rand(n, env.action_dim, generator=g, device="cuda") * 2 - 1)
        assert list(obs) == keys_both and env.single_observation_space.spaces["image_top"].contains(obs["image_top"][0].cpu().numpy())
        # the same poses through the numpy restatement
        rd = env._renderer
        poses = rd.poses.cpu().numpy().astype(np.float64)
        geoms, planes = rd.geoms.cpu().numpy().astype(np.float64), rd.planes.cpu().numpy().astype(np.float64)
        st = env.get_state()["qpos"].cpu().numpy()
        assert np.allclose(poses[:, 7, :3], st[:, 6:9], atol=1e-6)  # slot 7 = the (first) cube
        for k, cam in enumerate(("camera_front", "camera_top")):
            got = obs["image_" + cam.split("_")[1]].cpu().numpy()
            for e in range(n):
                ref, hit = reference_image(poses[e], geoms, planes, render.CAMERAS[cam].astype(np.float32).astype(np.float64), 240, 320)
                diff = np.abs(got[e].astype(int) - ref.astype(int)).max(-1)
                assert (diff <= 2).mean() >= 0.99 and diff.mean() < 0.5, (env_id, cam, e, (diff <= 2).mean(), diff.mean())
                assert (hit >= 0).sum() > 500  # the arm and the cube(s) are in view
        # images follow the state: moving the cube changes the top image, and only around the cube
        before = obs["image_top"].clone()
        q = env.get_state()["qpos"].clone()
        q[:, 6] += 0.05
        env.set_state(qpos=q)
        after = env._split(env._obs)["image_top"]
        changed = (before != after).any(-1).float().mean(dim=(1, 2))
        assert (changed > 0).all() and (changed < 0.05).all()
        env.close()
