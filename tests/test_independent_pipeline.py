"""A second, structurally different check of the collision pipeline: brute force in numpy / scipy against the oracle.

Nothing here shares code or data with the oracle or the kernels: the scene comes from tests/golden/independent/scene_push.npz,
written by tools/make_independent_scene.py with its own minimal MJCF / STL reader (hull vertices in the RAW mesh frames, not
the recentred ones of gym_lowcostrobot_b200/mjcf.py), forward kinematics are restated below, and the contact SET is found by
an exhaustive loop over every geom pair with exact tests -- no broadphase boxes, no separating-axis cache, no candidate /
job lists, no caps:

* pair filter = MuJoCo's published rule (geoms of bodies welded together never collide; neither do a body and its parent
  unless one of them is welded to the world; plus the <exclude> pairs of follower.xml);
* two hulls intersect iff no separating plane exists (a 4-variable linear programme over all hull vertices); for
  intersecting pairs the exact minimum penetration depth is the distance from the origin to the nearest facet of the
  Minkowski difference (Qhull);
* plane-hull and plane-box: a vertex below z = 0.

The oracle's contact list of the same state must contain exactly the intersecting pairs (pairs within 1e-6 of touching are
not judged), after a cold start AND after substeps that filled its separating-axis cache; every reported convex contact must
have a depth between the exact minimum and the hulls' overlap along its normal, and a position inside the overlap slab.
Reference: this is what `mujoco.mj_step` -> mj_collision does for the scene (reach_cube_env.py:276-277); SURVEY Appendix A.3.
"""
import numpy as np
import pytest
from scipy.optimize import linprog
from scipy.spatial import ConvexHull

from oracle.oracle import Oracle

from independent_scene import CORNERS, HI, LO, NB, NG, S, load_scene, world_boxes, world_geoms


def may_collide(b1, b2):
    """MuJoCo's body-pair filter for two arm bodies (body 0 = base_link is welded to the world)"""
    if b1 == b2:
        return False
    w1, w2 = (0 if b1 == 0 else b1), (0 if b2 == 0 else b2)  # weld ids: base_link shares the world's
    if w1 != 0 and w2 != 0 and (S["body_parent"][b1] == b2 or S["body_parent"][b2] == b1):
        return False
    return not any({b1, b2} == {int(a), int(b)} for a, b in S["exclude"])


def separation(A, B):
    """> 0: a separating plane exists (value = the gap it certifies); otherwise -(exact minimum penetration depth)"""
    cA, cB = A.mean(0), B.mean(0)
    if np.linalg.norm(cA - cB) > np.linalg.norm(A - cA, axis=1).max() + np.linalg.norm(B - cB, axis=1).max():
        return 1.0  # bounding spheres apart (a sound shortcut, nothing to do with the pipeline under test)
    # find n, d with n.a <= d - 1 (all a) and n.b >= d + 1 (all b)
    G = np.vstack([np.c_[A, -np.ones(len(A))], np.c_[-B, np.ones(len(B))]])
    res = linprog(np.zeros(4), A_ub=G, b_ub=-np.ones(len(G)), bounds=[(None, None)] * 4, method="highs")
    if res.status == 0:
        n = res.x[:3] / np.linalg.norm(res.x[:3])
        return (B @ n).min() - (A @ n).max()
    hull = ConvexHull((A[:, None, :] - B[None, :, :]).reshape(-1, 3))
    return hull.equations[:, 3].max()  # all offsets <= 0 when the origin is inside; the largest is minus the depth


def brute_force_pairs(qpos, Sc=S):
    """{(g1, g2): signed separation} over ALL geom pairs of the scene; geom ids: 0..19 meshes, 20 floor, 21.. boxes (the free cubes,
    then the static walls).  Walls and the floor belong to the world, like base_link which is welded to it: no pairs among those."""
    verts, _ = world_geoms(np.r_[qpos[:6], 0, 0, 0, 1, 0, 0, 0])
    boxes = world_boxes(Sc, qpos)
    ncube = int(Sc["ncube"])
    out = {}
    for g in range(NG):
        on_world = S["geom_body"][g] == 0
        if not on_world:
            out[(g, NG)] = verts[g][:, 2].min()
        for b, box in enumerate(boxes):
            if b < ncube or not on_world:
                out[(NG + 1 + b, g)] = separation(box, verts[g])
        for h in range(g + 1, NG):
            if may_collide(int(S["geom_body"][g]), int(S["geom_body"][h])):
                out[(g, h)] = separation(verts[g], verts[h])
    for c in range(ncube):
        out[(NG, NG + 1 + c)] = boxes[c][:, 2].min()
        for b in range(c + 1, len(boxes)):
            out[(NG + 1 + c, NG + 1 + b)] = separation(boxes[c], boxes[b])
    return out, verts, boxes


def oracle_contacts(o):
    return o.get("contacts").reshape(-1, 27)


def check_state(o, qpos, stats, Sc=S):
    """compare the oracle's contact list (already computed for qpos) with the brute force"""
    pairs, verts, boxes = brute_force_pairs(qpos, Sc)
    con = oracle_contacts(o)
    reported = {}
    for c in con:
        key = tuple(sorted((int(c[14]), int(c[15]))))
        reported.setdefault(key, []).append(c)
    geom = lambda g: boxes[g - NG - 1] if g > NG else verts[g]
    for (g1, g2), sep in pairs.items():
        key = tuple(sorted((g1, g2)))
        if abs(sep) < 1e-6:
            stats["borderline"] += 1
            continue
        if sep < 0:
            assert key in reported, f"oracle misses the penetrating pair {key} (depth {-sep:.3e})"
            stats["hits"] += 1
        else:
            assert key not in reported, f"oracle reports a contact for the separated pair {key} (gap {sep:.3e})"
    assert set(reported) <= {tuple(sorted(k)) for k in pairs}, "contact between geoms the filter excludes"
    for key, cs in reported.items():
        if NG in key:
            continue  # plane contacts: test_floor_contacts_are_hull_vertices_below_the_plane
        for c in cs:
            A, B, n, depth = geom(int(c[14])), geom(int(c[15])), c[3:6], -c[12]
            exact = -pairs[(int(c[14]), int(c[15]))] if (int(c[14]), int(c[15])) in pairs else -pairs[(int(c[15]), int(c[14]))]
            overlap = (A @ n).max() - (B @ n).min()
            assert abs(np.linalg.norm(n) - 1) < 1e-9
            if len(cs) == 1 and min(key) < NG:  # one MPR contact per convex pair (box-box pairs: SAT + clipping, several points)
                assert exact - 1e-5 <= depth <= overlap + 1e-6, (key, exact, depth, overlap)
            assert (B @ n).min() - 1e-6 <= c[0:3] @ n <= (A @ n).max() + 1e-6
            stats["depths"] += 1


def test_own_kinematics_reproduce_the_oracles_frames():
    """sanity of the independent reader itself: hull vertices in the raw mesh frames + own FK land where the oracle's
    recentred meshes + its kinematics put them (same hull sizes in the same geom order, matching world bounding boxes)"""
    from gym_lowcostrobot_b200 import model

    m = model.load_compiled("push")
    assert list(np.diff(S["hull_adr"])) == list(m["mesh_vertnum"]) and list(S["geom_body"] ) == [int(b) for b in m["mesh_body"]]
    rng = np.random.default_rng(0)
    for _ in range(5):
        qpos = np.r_[rng.uniform(LO, HI), 0.0, 0.2, 0.015, 1, 0, 0, 0]
        o = Oracle("push")
        o.set_state(qpos=qpos, qvel=np.zeros(12), ctrl=qpos[:6])
        o.forward()
        xpos, xmat = o.get("xpos").reshape(-1, 3), o.get("xmat").reshape(-1, 3, 3)
        verts, _ = world_geoms(qpos)
        for g in range(NG):
            b = int(m["mesh_body"][g])
            w = xpos[b] + m["verts"][m["mesh_vertadr"][g]: m["mesh_vertadr"][g] + m["mesh_vertnum"][g]] @ xmat[b].T
            assert np.abs(w.min(0) - verts[g].min(0)).max() < 1e-6 and np.abs(w.max(0) - verts[g].max(0)).max() < 1e-6


def test_exhaustive_pair_loop_agrees_with_the_oracles_contact_set():
    rng = np.random.default_rng(5)
    stats = dict(hits=0, depths=0, borderline=0)
    kinds = set()
    for trial in range(14):
        qpos = np.zeros(13)
        qpos[9] = 1
        qpos[:6] = rng.uniform(LO, HI)
        if trial % 3 == 0:  # folded arm: self collisions and links in the floor
            qpos[1], qpos[2] = rng.uniform(0.8, 1.22), rng.uniform(1.0, 1.74)
        verts, _ = world_geoms(qpos)
        g = rng.integers(3, NG)
        qpos[6:9] = verts[g].mean(0) + rng.uniform(-0.02, 0.02, 3) if trial % 2 else [rng.uniform(-0.1, 0.1), rng.uniform(0.1, 0.25), rng.uniform(0.0, 0.03)]
        q = rng.normal(size=4)
        qpos[9:13] = q / np.linalg.norm(q)
        o = Oracle("push")
        o.set_state(qpos=qpos, qvel=np.zeros(12), ctrl=qpos[:6], warm=np.zeros(12))
        o.forward()  # cold start: empty separating-axis cache
        check_state(o, qpos, stats)
        for c in oracle_contacts(o):
            a, b = int(c[14]), int(c[15])
            kinds.add("floor" if NG in (a, b) else "cube" if max(a, b) > NG else "self")
        if trial % 2 == 0:  # warm: a few substeps fill the cache, then the contact list of the state reached
            o.set_state(qvel=rng.normal(scale=0.3, size=12))
            o.substep(3)
            q1 = o.get_state()["qpos"]
            o.forward()
            check_state(o, q1, stats)
    print(stats, kinds)
    assert stats["hits"] >= 40 and stats["depths"] >= 15 and kinds == {"floor", "cube", "self"}, (stats, kinds)


@pytest.mark.parametrize("task", ["stack", "push_loop"])
def test_exhaustive_pair_loop_on_the_other_scene_classes(task):
    """the same check on the two other scene classes of the kernels: two free cubes (cube-cube by SAT + clipping), and one cube
    inside four static rails (wall-cube, wall-mesh)"""
    Sc = load_scene(task)
    rng = np.random.default_rng(9)
    stats = dict(hits=0, depths=0, borderline=0)
    seen = set()
    ncube, nq = int(Sc["ncube"]), 6 + 7 * int(Sc["ncube"])
    for trial in range(12):
        qpos = np.zeros(nq)
        qpos[:6] = rng.uniform(LO, HI)
        if task == "push_loop" and trial % 2:
            qpos[1], qpos[2] = rng.uniform(0.6, 1.1), rng.uniform(0.8, 1.5)  # gripper lowered towards the rails
        verts, _ = world_geoms(np.r_[qpos[:6], 0, 0, 0, 1, 0, 0, 0])
        for c in range(ncube):
            p = qpos[6 + 7 * c: 13 + 7 * c]
            if task == "push_loop":  # on / inside a rail or free in the pen
                w = rng.integers(0, 4)
                p[:3] = Sc["wall_pos"][w] + rng.uniform(-1, 1, 3) * (Sc["box_half"][1 + w] + 0.012) + [0, 0, 0.008] if trial % 3 else [rng.uniform(-0.1, 0.1), rng.uniform(0.1, 0.17), 0.015]
            elif c == 0:
                p[:3] = verts[rng.integers(3, NG)].mean(0) + rng.uniform(-0.02, 0.02, 3) if trial % 2 else [rng.uniform(-0.1, 0.1), rng.uniform(0.1, 0.25), rng.uniform(0.0, 0.03)]
            else:  # the second cube on / inside / beside the first
                p[:3] = qpos[6:9] + rng.uniform(-0.025, 0.025, 3) + [0, 0, 0.02 * (trial % 3)]
            q = rng.normal(size=4)
            p[3:7] = q / np.linalg.norm(q) if trial % 4 else [1, 0, 0, 0]
        o = Oracle(task)
        o.set_state(qpos=qpos, qvel=np.zeros(o.nv), ctrl=qpos[:6], warm=np.zeros(o.nv))
        o.forward()
        check_state(o, qpos, stats, Sc)
        for c in oracle_contacts(o):
            a, b = sorted((int(c[14]), int(c[15])))
            seen.add("cube-cube" if a > NG and b <= NG + ncube else "wall-cube" if a > NG else "wall-mesh" if b > NG + ncube else "other")
        if trial % 2 == 0:
            o.set_state(qvel=rng.normal(scale=0.3, size=o.nv))
            o.substep(3)
            q1 = o.get_state()["qpos"]
            o.forward()
            check_state(o, q1, stats, Sc)
    want = {"cube-cube"} if task == "stack" else {"wall-cube", "wall-mesh"}
    print(task, stats, seen)
    assert stats["hits"] >= 40 and want <= seen, (stats, seen)
