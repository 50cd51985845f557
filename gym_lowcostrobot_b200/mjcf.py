"""Host-side model compiler: MJCF subset + binary STL -> flat model constants.

The reference loads its scenes with ``mujoco.MjModel.from_xml_path`` once per env
(reference ``gym_lowcostrobot/envs/reach_cube_env.py:89``, ``push_cube_env.py:92``,
``lift_cube_env.py:90``, ``pick_place_cube_env.py:93``, ``stack_two_cubes_env.py:90``).
MuJoCo is not available in this image, so this module re-implements the part of the
MJCF compiler those scenes need (``assets/low_cost_robot_6dof/follower.xml`` + one
scene file): ``<include>`` in place, nested ``<default>`` classes with ``childclass``,
``<option>`` merge in document order, ``<compiler angle/meshdir>``, mesh assets
(convex hulls via Qhull), the body tree with ``<inertial>``, hinge/free joints,
geoms, one site, ``<contact><exclude>`` and ``<position>`` actuators.

The output is a plain ``dict`` of numpy arrays (see ``compile_model``) that
``model.py`` packs into the ``LcrModel`` C struct of ``include/lcr_model.h``.
The device never sees XML.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

TASK_IDS = {"reach": 0, "push": 1, "lift": 2, "pick_place": 3, "stack": 4, "push_loop": 5}
TASK_XML = {
    "reach": "reach_cube.xml",
    "push": "push_cube.xml",
    "lift": "lift_cube.xml",
    "pick_place": "pick_place_cube.xml",
    "stack": "stack_two_cubes.xml",
    "push_loop": "push_cube_loop.xml",
}

# MuJoCo built-in element defaults (public MJCF reference documentation).
_GEOM_BUILTIN = dict(
    type="sphere", contype="1", conaffinity="1", condim="3", priority="0",
    friction="1 0.005 0.0001", solmix="1", solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2",
    margin="0", gap="0", pos="0 0 0", quat="1 0 0 0",
)
_JOINT_BUILTIN = dict(
    type="hinge", pos="0 0 0", axis="0 0 1", armature="0", damping="0",
    solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2", margin="0",
)
_OPTION_BUILTIN = dict(
    timestep="0.002", impratio="1", gravity="0 0 -9.81", tolerance="1e-8", ls_tolerance="0.01",
    iterations="100", ls_iterations="50", integrator="Euler", cone="pyramidal", solver="Newton",
)


def _floats(s, n=None):
    v = np.array([float(x) for x in str(s).split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


def _expand_includes(elem, base_dir):
    """MJCF ``<include>``: the included file's top-level children replace the element in place."""
    out = []
    for child in list(elem):
        if child.tag == "include":
            sub = ET.parse(os.path.join(base_dir, child.attrib["file"])).getroot()
            _expand_includes(sub, base_dir)
            out.extend(list(sub))
        else:
            _expand_includes(child, base_dir)
            out.append(child)
    for c in list(elem):
        elem.remove(c)
    for c in out:
        elem.append(c)


class _Defaults:
    """Nested default classes: class name -> {element tag -> attribute dict}, inherited from the parent."""

    def __init__(self):
        self.classes = {"main": {}}

    def load(self, default_elem, parent="main", top=True):
        name = "main" if top else default_elem.attrib["class"]
        if name not in self.classes:
            self.classes[name] = {k: dict(v) for k, v in self.classes[parent].items()}
        for child in default_elem:
            if child.tag == "default":
                self.load(child, parent=name, top=False)
            else:
                self.classes[name].setdefault(child.tag, {}).update(child.attrib)
        # children classes were copied before later siblings were read; MJCF files in scope
        # declare element defaults before nested classes, which the copy above relies on.

    def resolve(self, tag, elem, childclass, builtin):
        cls = elem.attrib.get("class", childclass or "main")
        if cls not in self.classes:
            raise ValueError(f"unknown default class {cls!r}")
        a = dict(builtin)
        a.update(self.classes[cls].get(tag, {}))
        a.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return a


def _quat_normalize(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.linalg.norm(q)


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def load_stl_triangles(path):
    """Triangles [n,3,3] of a binary STL (80-byte header, uint32 count, 50-byte records)."""
    raw = open(path, "rb").read()
    (n,) = struct.unpack("<I", raw[80:84])
    if len(raw) < 84 + 50 * n:
        raise ValueError(f"{path}: not a binary STL")
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    tri = np.frombuffer(raw, dtype=rec, count=n, offset=84)
    return tri["v"].astype(np.float64)


def load_stl_vertices(path):
    """Unique vertices of a binary STL."""
    return np.unique(load_stl_triangles(path).reshape(-1, 3), axis=0)


def mesh_volume_centroid(tri):
    """Centroid of the volume enclosed by a triangle mesh (signed tetrahedra against the vertex mean)."""
    ref = tri.reshape(-1, 3).mean(0)
    a, b, c = tri[:, 0] - ref, tri[:, 1] - ref, tri[:, 2] - ref
    vol = np.einsum("ij,ij->i", a, np.cross(b, c)) / 6.0
    cen = (a + b + c) / 4.0
    if abs(vol.sum()) < 1e-18:
        return ref
    return ref + (vol[:, None] * cen).sum(0) / vol.sum()


def convex_hull_vertices(points):
    """Vertices of the convex hull (MuJoCo collides mesh geoms as their convex hulls)."""
    from scipy.spatial import ConvexHull

    hull = ConvexHull(points)
    return points[np.sort(hull.vertices)]


def compile_model(assets_dir, task):
    """Parse ``<assets_dir>/<scene>.xml`` and return the model constants as a dict of numpy arrays."""
    xml = os.path.join(assets_dir, TASK_XML[task])
    root = ET.parse(xml).getroot()
    _expand_includes(root, assets_dir)

    # ---- compiler / option (document order, later attributes override) ----
    meshdir = ""
    angle = "degree"
    for c in root.findall("compiler"):
        meshdir = c.attrib.get("meshdir", meshdir)
        angle = c.attrib.get("angle", angle)
    if angle != "radian":
        raise NotImplementedError("only angle='radian' models are supported")
    opt = dict(_OPTION_BUILTIN)
    for o in root.findall("option"):
        opt.update(o.attrib)
    if opt["integrator"] != "implicitfast" or opt["cone"] != "elliptic" or opt["solver"] != "Newton":
        raise NotImplementedError(f"unsupported option set {opt}")

    defaults = _Defaults()
    for d in root.findall("default"):
        defaults.load(d)

    mesh_files = {}
    for a in root.findall("asset"):
        for m in a.findall("mesh"):
            mesh_files[m.attrib["name"]] = os.path.join(assets_dir, meshdir, m.attrib["file"])

    # ---- walk the body tree ----
    arm_bodies, cubes, meshes, walls = [], [], [], []
    goals = {}
    floor = None
    site = None
    joint_names = []
    body_names = {}

    def geom_params(a):
        fr = _floats(a["friction"])
        fr = np.concatenate([fr, _floats(_GEOM_BUILTIN["friction"])[fr.size:]])
        return dict(
            condim=int(a["condim"]), priority=int(a["priority"]), friction=fr,
            solref=_floats(a["solref"], 2), solimp=np.concatenate(
                [_floats(a["solimp"]), _floats(_GEOM_BUILTIN["solimp"])[_floats(a["solimp"]).size:]]),
            solmix=float(a["solmix"]), margin=float(a["margin"]), gap=float(a["gap"]),
            contype=int(a["contype"]), conaffinity=int(a["conaffinity"]),
        )

    def world_geom(g):
        """Colliding geom of the world body (or of a jointless static body at the world origin): the z = 0 floor plane or
        an axis-aligned static box (the rails of push_cube_loop.xml:45-48); non-colliding boxes named goal_region_* are
        the goal regions push_cube_loop_env.py:127-133 reads from the model."""
        nonlocal floor
        ga = defaults.resolve("geom", g, None, _GEOM_BUILTIN)
        p = geom_params(ga)
        name = ga.get("name", "")
        if not (p["contype"] or p["conaffinity"]):
            if name.startswith("goal_region_") and ga["type"] == "box":
                goals[name] = (_floats(ga["pos"], 3), _floats(ga["size"], 3))
            return  # target_region: visual only (push_cube.xml:35)
        if ga["type"] == "plane":
            if np.any(_floats(ga["pos"], 3) != 0) or np.any(_quat_normalize(_floats(ga["quat"], 4)) != [1, 0, 0, 0]):
                raise NotImplementedError("the only plane supported is z = 0")
            floor = p
        elif ga["type"] == "box":
            if np.any(_quat_normalize(_floats(ga["quat"], 4)) != [1, 0, 0, 0]) or "euler" in ga:
                raise NotImplementedError("static boxes must be axis aligned")
            p.update(pos=_floats(ga["pos"], 3), size=_floats(ga["size"], 3), name=name)
            walls.append(p)
        else:
            raise NotImplementedError("colliding world geoms: the z = 0 plane and axis-aligned boxes")

    def walk(body, parent_idx, childclass, depth):
        nonlocal floor, site
        if (parent_idx is None and not body.findall("joint") and not body.findall("freejoint") and not body.findall("body")
                and body.find("inertial") is None and all("mesh" not in g.attrib for g in body.findall("geom"))):
            # static body welded to the world holding world geometry (<body name="floor">, push_cube_loop.xml:24-26)
            if np.any(_floats(body.attrib.get("pos", "0 0 0"), 3) != 0) or "quat" in body.attrib or "euler" in body.attrib:
                raise NotImplementedError("static bodies must sit at the world origin")
            for g in body.findall("geom"):
                world_geom(g)
            return
        childclass = body.attrib.get("childclass", childclass)
        name = body.attrib.get("name", "")
        joints = body.findall("joint")
        free = body.findall("freejoint")
        rec = dict(
            name=name,
            pos=_floats(body.attrib.get("pos", "0 0 0"), 3),
            quat=_quat_normalize(_floats(body.attrib.get("quat", "1 0 0 0"), 4)),
            parent=parent_idx,
        )
        inertial = body.find("inertial")
        if inertial is not None:
            rec.update(
                ipos=_floats(inertial.attrib.get("pos", "0 0 0"), 3),
                iquat=_quat_normalize(_floats(inertial.attrib.get("quat", "1 0 0 0"), 4)),
                mass=float(inertial.attrib["mass"]),
                inertia=_floats(inertial.attrib["diaginertia"], 3),
            )
        else:
            rec.update(ipos=np.zeros(3), iquat=np.array([1.0, 0, 0, 0]), mass=0.0, inertia=np.zeros(3))
        if free:
            kind = "cube"
            idx = len(cubes)
            cubes.append(rec)
        else:
            kind = "arm"
            idx = len(arm_bodies)
            arm_bodies.append(rec)
            if len(joints) > 1:
                raise NotImplementedError("one joint per body")
            if joints:
                ja = defaults.resolve("joint", joints[0], childclass, _JOINT_BUILTIN)
                if ja["type"] != "hinge" or np.any(_floats(ja["pos"], 3) != 0):
                    raise NotImplementedError("only hinge joints anchored at the body origin")
                ax = _floats(ja["axis"], 3)
                rec["joint"] = dict(
                    name=ja.get("name", ""), axis=ax / np.linalg.norm(ax), range=_floats(ja["range"], 2),
                    armature=float(ja["armature"]), damping=float(ja["damping"]),
                    frcrange=_floats(ja.get("actuatorfrcrange", "0 0"), 2),
                    solref=_floats(ja["solreflimit"], 2), solimp=_floats(ja["solimplimit"], 5),
                )
                joint_names.append(ja.get("name", ""))
            elif parent_idx is not None:
                raise NotImplementedError("fixed child bodies are not supported")
        body_names[name] = (kind, idx)
        for g in body.findall("geom"):
            ga = defaults.resolve("geom", g, childclass if kind == "arm" else None, _GEOM_BUILTIN)
            gtype = "mesh" if "mesh" in ga else ga["type"]
            p = geom_params(ga)
            if not (p["contype"] or p["conaffinity"]):
                continue
            if np.any(_floats(ga["pos"], 3) != 0) or np.any(_quat_normalize(_floats(ga["quat"], 4)) != [1, 0, 0, 0]):
                raise NotImplementedError("geom frames must coincide with the body frame")
            if gtype == "mesh" and kind == "arm":
                p.update(body=idx, mesh=ga["mesh"])
                meshes.append(p)
            elif gtype == "box" and kind == "cube":
                p.update(size=_floats(ga["size"], 3))
                rec["geom"] = p
            else:
                raise NotImplementedError(f"geom type {gtype} on {kind} body")
        for s in body.findall("site"):
            if s.attrib.get("name") == "end_effector_site":
                site = dict(body=idx, pos=_floats(s.attrib.get("pos", "0 0 0"), 3))
        for child in body.findall("body"):
            if kind != "arm":
                raise NotImplementedError("children of free bodies")
            walk(child, idx, childclass, depth + 1)

    for wb in root.findall("worldbody"):
        for g in wb.findall("geom"):
            world_geom(g)
        for b in wb.findall("body"):
            walk(b, None, None, 0)

    na = len(arm_bodies)
    if na != 7 or len(joint_names) != 6 or site is None or floor is None:
        raise NotImplementedError("expected base_link + 6 hinged links, a floor and the ee site")
    for k, b in enumerate(arm_bodies):
        if b["parent"] != (k - 1 if k else None):
            raise NotImplementedError("arm must be a serial chain")
    ncube = len(cubes)
    if ncube not in (1, 2):
        raise NotImplementedError("1 or 2 free cubes")

    # ---- actuators: <position> on each hinge, in joint order ----
    act = []
    for a in root.findall("actuator"):
        for p in a:
            if p.tag != "position":
                raise NotImplementedError(p.tag)
            pa = defaults.resolve("position", p, None, dict(kp="1", kv="0", inheritrange="0"))
            act.append(pa)
    if [a["joint"] for a in act] != joint_names:
        raise NotImplementedError("one position actuator per joint, in joint order")

    excludes = set()
    for c in root.findall("contact"):
        for e in c.findall("exclude"):
            excludes.add(frozenset((body_names[e.attrib["body1"]][1], body_names[e.attrib["body2"]][1])))

    # ---- meshes -> hull vertex pool (body frame) ----
    verts, vertadr, vertnum = [], [], []
    hull_cache = {}
    for m in meshes:
        if m["mesh"] not in hull_cache:
            tri = load_stl_triangles(mesh_files[m["mesh"]])
            hull_cache[m["mesh"]] = (convex_hull_vertices(np.unique(tri.reshape(-1, 3), axis=0)), mesh_volume_centroid(tri))
        hv, m["com"] = hull_cache[m["mesh"]]
        vertadr.append(sum(vertnum))
        vertnum.append(len(hv))
        verts.append(hv)
        lo, hi = hv.min(0), hv.max(0)
        m["center"] = 0.5 * (lo + hi)
        m["half"] = 0.5 * (hi - lo)
        m["rbound"] = float(np.sqrt(((hv - m["center"]) ** 2).sum(1).max()))
    verts = np.concatenate(verts, 0)

    # ---- candidate mesh-mesh pairs (MuJoCo body-pair filter: same weld body, parent-child unless
    # the parent is welded to the world, explicit excludes).  base_link (arm body 0) is welded to
    # the world; its weld parent is the world too. ----
    def weld(b):
        return 0 if b == 0 else b  # arm body index doubles as weld id; 0 == world

    pairs = []
    for i, gi in enumerate(meshes):
        for j in range(i + 1, len(meshes)):
            gj = meshes[j]
            b1, b2 = gi["body"], gj["body"]
            if weld(b1) == weld(b2):
                continue
            w1, w2 = weld(b1), weld(b2)
            p1, p2 = weld(max(b1 - 1, 0)), weld(max(b2 - 1, 0))
            if w1 != 0 and w2 != 0 and (w1 == p2 or w2 == p1):
                continue
            if frozenset((b1, b2)) in excludes:
                continue
            if not ((gi["contype"] & gj["conaffinity"]) or (gj["contype"] & gi["conaffinity"])):
                continue
            pairs.append((i, j))

    out = dict(
        task=np.int32(TASK_IDS[task]), ncube=np.int32(ncube),
        timestep=float(opt["timestep"]), impratio=float(opt["impratio"]), gravity=_floats(opt["gravity"], 3),
        tolerance=float(opt["tolerance"]), ls_tolerance=float(opt["ls_tolerance"]),
        iterations=np.int32(opt["iterations"]), ls_iterations=np.int32(opt["ls_iterations"]),
        body_pos=np.stack([b["pos"] for b in arm_bodies]), body_quat=np.stack([b["quat"] for b in arm_bodies]),
        body_ipos=np.stack([b["ipos"] for b in arm_bodies]), body_iquat=np.stack([b["iquat"] for b in arm_bodies]),
        body_mass=np.array([b["mass"] for b in arm_bodies]), body_inertia=np.stack([b["inertia"] for b in arm_bodies]),
        jnt_axis=np.stack([b["joint"]["axis"] for b in arm_bodies[1:]]),
        jnt_range=np.stack([b["joint"]["range"] for b in arm_bodies[1:]]),
        jnt_armature=np.array([b["joint"]["armature"] for b in arm_bodies[1:]]),
        jnt_damping=np.array([b["joint"]["damping"] for b in arm_bodies[1:]]),
        jnt_frcrange=np.stack([b["joint"]["frcrange"] for b in arm_bodies[1:]]),
        jnt_solref=np.stack([b["joint"]["solref"] for b in arm_bodies[1:]]),
        jnt_solimp=np.stack([b["joint"]["solimp"] for b in arm_bodies[1:]]),
        act_kp=np.array([float(a["kp"]) for a in act]), act_kv=np.array([float(a["kv"]) for a in act]),
        site_body=np.int32(site["body"]), site_pos=site["pos"],
        cube_mass=np.array([c["mass"] for c in cubes]), cube_inertia=np.stack([c["inertia"] for c in cubes]),
        cube_size=np.stack([c["geom"]["size"] for c in cubes]),
        cube_pos0=np.stack([c["pos"] for c in cubes]),
        verts=verts, mesh_vertadr=np.array(vertadr, np.int32), mesh_vertnum=np.array(vertnum, np.int32),
        mesh_body=np.array([m["body"] for m in meshes], np.int32),
        mesh_center=np.stack([m["center"] for m in meshes]), mesh_half=np.stack([m["half"] for m in meshes]),
        mesh_rbound=np.array([m["rbound"] for m in meshes]), mesh_com=np.stack([m["com"] for m in meshes]),
        pair_g1=np.array([p[0] for p in pairs], np.int32), pair_g2=np.array([p[1] for p in pairs], np.int32),
        mesh_names=np.array([m["mesh"] for m in meshes]),
    )
    # ctrlrange: inheritrange=1 copies the joint range (follower.xml:8)
    ctrl = []
    for a, b in zip(act, arm_bodies[1:]):
        if float(a["inheritrange"]) != 1.0:
            raise NotImplementedError("position actuators must use inheritrange=1")
        ctrl.append(b["joint"]["range"])
    out["act_ctrlrange"] = np.stack(ctrl)
    for c in cubes:
        if np.any(c["ipos"] != 0) or np.any(c["iquat"] != [1, 0, 0, 0]) or len(set(c["inertia"])) != 1:
            raise NotImplementedError("cube inertia must be isotropic and centred")

    if walls:
        out["wall_pos"] = np.stack([w["pos"] for w in walls])
        out["wall_size"] = np.stack([w["size"] for w in walls])
        out["wall_names"] = np.array([w["name"] for w in walls])
    if goals:
        if sorted(goals) != ["goal_region_1", "goal_region_2"]:
            raise NotImplementedError("expected goal_region_1 and goal_region_2")
        out["goal_center"] = np.stack([goals["goal_region_1"][0], goals["goal_region_2"][0]])
        out["goal_size"] = goals["goal_region_1"][1]
    if task == "push_loop" and (len(walls) == 0 or not goals or ncube != 1):
        raise NotImplementedError("push_loop needs one cube, the rails and the two goal regions")

    # geom parameter table: rows 0..nmesh-1 arm meshes, then floor, cube0, cube1, walls
    geoms = meshes + [floor] + [c["geom"] for c in cubes] + walls
    out["geom_condim"] = np.array([g["condim"] for g in geoms], np.int32)
    out["geom_priority"] = np.array([g["priority"] for g in geoms], np.int32)
    out["geom_friction"] = np.stack([g["friction"] for g in geoms])
    out["geom_solref"] = np.stack([g["solref"] for g in geoms])
    out["geom_solimp"] = np.stack([g["solimp"] for g in geoms])
    out["geom_solmix"] = np.array([g["solmix"] for g in geoms])
    for g in geoms:
        if g["margin"] != 0 or g["gap"] != 0:
            raise NotImplementedError("geom margin/gap")

    _set_const(out)
    return out


# --------------------------------------------------------------------------------------
# qpos0-dependent constants (MuJoCo mj_setConst): invweight0, meaninertia.
# --------------------------------------------------------------------------------------
def arm_kinematics(m, q):
    """World poses of the 7 arm bodies: (xpos[7,3], xmat[7,3,3], axis_world[6,3])."""
    xpos = np.zeros((7, 3))
    xquat = np.zeros((7, 4))
    axis = np.zeros((6, 3))
    p, qt = np.zeros(3), np.array([1.0, 0, 0, 0])
    for b in range(7):
        p = p + quat_to_mat(qt) @ m["body_pos"][b]
        qt = quat_mul(qt, m["body_quat"][b])
        if b >= 1:
            a = m["jnt_axis"][b - 1]
            axis[b - 1] = quat_to_mat(qt) @ a
            h = 0.5 * q[b - 1]
            qt = quat_mul(qt, np.concatenate([[np.cos(h)], np.sin(h) * a]))
        xpos[b], xquat[b] = p, qt
    xmat = np.stack([quat_to_mat(x) for x in xquat])
    return xpos, xmat, axis


def arm_mass_matrix(m, q, armature=True):
    """Joint-space inertia of the arm, M = sum_b m Jv^T Jv + Jw^T I Jw (+ armature)."""
    xpos, xmat, axis = arm_kinematics(m, q)
    M = np.zeros((6, 6))
    jacs = []
    for b in range(1, 7):
        com = xpos[b] + xmat[b] @ m["body_ipos"][b]
        Ri = xmat[b] @ quat_to_mat(m["body_iquat"][b])
        Iw = Ri @ np.diag(m["body_inertia"][b]) @ Ri.T
        Jv, Jw = np.zeros((3, 6)), np.zeros((3, 6))
        for j in range(b):  # joint j sits on body j+1 <= b
            Jw[:, j] = axis[j]
            Jv[:, j] = np.cross(axis[j], com - xpos[j + 1])
        M += m["body_mass"][b] * Jv.T @ Jv + Jw.T @ Iw @ Jw
        jacs.append((Jv, Jw))
    if armature:
        M += np.diag(m["jnt_armature"])
    return M, jacs


def _set_const(m):
    q0 = np.zeros(6)
    M, jacs = arm_mass_matrix(m, q0)
    Minv = np.linalg.inv(M)
    m["dof_invweight0"] = np.diag(Minv).copy()
    bw = np.zeros((7, 2))
    for b in range(1, 7):
        Jv, Jw = jacs[b - 1]
        bw[b, 0] = np.trace(Jv @ Minv @ Jv.T) / 3.0
        bw[b, 1] = np.trace(Jw @ Minv @ Jw.T) / 3.0
    m["body_invweight0"] = bw
    m["cube_invweight0"] = np.stack([1.0 / m["cube_mass"], 1.0 / m["cube_inertia"][:, 0]], 1)
    diag = list(np.diag(M))
    for c in range(int(m["ncube"])):
        diag += [m["cube_mass"][c]] * 3 + list(m["cube_inertia"][c])
    m["meaninertia"] = float(np.mean(diag))
