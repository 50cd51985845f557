"""Fixture for tests/test_independent_pipeline.py: the scenes of push_cube.xml, stack_two_cubes.xml and
push_cube_loop.xml read by a SEPARATE minimal reader.

Deliberately shares nothing with gym_lowcostrobot_b200/mjcf.py (the model compiler that the oracle and the CUDA path both
use): plain ElementTree walk of follower.xml + push_cube.xml, raw binary STL triangles, scipy's Qhull for the hull vertex
sets, vertices kept in the RAW mesh frame (MuJoCo recentres meshes on their centre of mass and compensates in the geom
pose; an independent reader has no reason to).  Usage (in the build container, where /root/reference exists):

    python tools/make_independent_scene.py [/root/reference/gym_lowcostrobot/assets/low_cost_robot_6dof]

writes tests/golden/independent/scene_{push,stack,push_loop}.npz.
"""
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial import ConvexHull

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_stl(path):
    raw = open(path, "rb").read()
    (ntri,) = struct.unpack_from("<I", raw, 80)
    assert len(raw) == 84 + 50 * ntri, "binary STL expected"
    tri = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=ntri, offset=84)
    return tri["v"].reshape(-1, 3).astype(np.float64)


def vec(s, n):
    v = [float(x) for x in s.split()]
    assert len(v) == n
    return np.array(v)


SCENES = {"push": "push_cube.xml", "stack": "stack_two_cubes.xml", "push_loop": "push_cube_loop.xml"}


def main(assets, task="push"):
    arm = ET.parse(os.path.join(assets, "follower.xml")).getroot()
    scene = ET.parse(os.path.join(assets, SCENES[task])).getroot()
    meshdir = os.path.join(assets, arm.find("compiler").get("meshdir"))
    files = {m.get("name"): m.get("file") for m in arm.find("asset").findall("mesh")}
    on_disk = {f.lower(): f for f in os.listdir(meshdir)}

    bodies, geoms = [], []  # bodies: name, parent index, pos, quat, joint axis (or zeros); geoms: body index, mesh name

    # default classes (only what the contact / joint / actuator parameters need): class -> parent class, per-element attributes
    classes = {}

    def read_defaults(elem, parent):
        name = elem.get("class", "main")
        classes[name] = (parent, {child.tag: dict(child.attrib) for child in elem if child.tag != "default"})
        for child in elem.findall("default"):
            read_defaults(child, name)

    read_defaults(arm.find("default"), None)

    def resolved(tag, cls, own):
        chain = []
        while cls is not None:
            chain.append(classes[cls][1].get(tag, {}))
            cls = classes[cls][0]
        out = {}
        for attrs in reversed(chain):
            out.update(attrs)
        out.update(own)
        return out

    def contact_params(attrs):
        """friction[3], condim, priority, solref[2], solimp[5], solmix with MuJoCo's documented defaults; a shorter attribute
        sets the leading components only"""
        def lead(default, text):
            v = list(default)
            if text is not None:
                given = [float(x) for x in text.split()]
                v[:len(given)] = given
            return v
        return (lead([1, 0.005, 0.0001], attrs.get("friction")) + [float(attrs.get("condim", 3)), float(attrs.get("priority", 0))] +
                lead([0.02, 1], attrs.get("solref")) + lead([0.9, 0.95, 0.001, 0.5, 2], attrs.get("solimp")) + [float(attrs.get("solmix", 1))])

    geom_par, inertial, joints = [], [], []

    def walk(elem, parent):
        idx = len(bodies)
        j = elem.find("joint")
        bodies.append(dict(name=elem.get("name"), parent=parent, pos=vec(elem.get("pos", "0 0 0"), 3), quat=vec(elem.get("quat", "1 0 0 0"), 4),
                           axis=vec(j.get("axis"), 3) if j is not None else np.zeros(3)))
        childclass = elem.get("childclass") or getattr(walk, "childclass", None)
        walk.childclass = childclass
        ine = elem.find("inertial")
        inertial.append(np.r_[vec(ine.get("pos"), 3), vec(ine.get("quat"), 4), float(ine.get("mass")), vec(ine.get("diaginertia"), 3)]
                        if ine is not None else np.zeros(11))
        if j is not None:
            ja = resolved("joint", j.get("class", childclass), dict(j.attrib))
            joints.append(np.r_[float(ja["armature"]), float(ja["damping"]), vec(ja["actuatorfrcrange"], 2), vec(ja["range"], 2)])
        for g in elem.findall("geom"):
            assert g.get("pos") is None and g.get("quat") is None  # the arm's mesh geoms sit in the body frame
            geoms.append((idx, g.get("mesh")))
            geom_par.append(contact_params(resolved("geom", g.get("class", childclass), dict(g.attrib))))
        for child in elem.findall("body"):
            walk(child, idx)

    walk(arm.find("worldbody").find("body"), -1)
    hull_pts, hull_adr = [], [0]
    for _, mesh in geoms:
        v = read_stl(os.path.join(meshdir, on_disk[files[mesh].lower()]))
        v = np.unique(v, axis=0)
        hv = v[ConvexHull(v).vertices]
        hull_pts.append(hv)
        hull_adr.append(hull_adr[-1] + len(hv))
    excl = [(e.get("body1"), e.get("body2")) for e in arm.find("contact").findall("exclude")]
    names = [b["name"] for b in bodies]
    # free bodies (one box geom each) = the cubes; box geoms directly in the worldbody that take part in collisions = static walls
    cubes = [b for b in scene.find("worldbody").findall("body") if b.find("freejoint") is not None]
    walls = [g for g in scene.find("worldbody").findall("geom") if g.get("type") == "box" and g.get("contype") != "0"]
    cube = cubes[0]
    floor = next(g for g in scene.find("worldbody").iter("geom") if g.get("name") == "floor")  # (PushCubeLoop wraps it in a jointless body)
    geom_par.append(contact_params(dict(floor.attrib)))              # geom 20: the floor plane
    for b in cubes:                                                   # geoms 21..: the cubes, then the static walls
        geom_par.append(contact_params(dict(b.find("geom").attrib)))
    for g in walls:
        geom_par.append(contact_params(dict(g.attrib)))
    cine = cube.find("inertial")
    act = resolved("position", "follower", {})
    # <option>: the include is expanded in place, so follower.xml's <option> comes after push_cube.xml's and wins where both set
    # an attribute (impratio 100 over 10); attributes only one of them sets are kept
    opt = dict(scene.find("option").attrib)
    opt.update(arm.find("option").attrib)
    assert opt["cone"] == "elliptic" and opt["integrator"] == "implicitfast"
    out = os.path.join(ROOT, "tests", "golden", "independent", f"scene_{task}.npz")
    np.savez_compressed(
        out, body_parent=np.array([b["parent"] for b in bodies]), body_pos=np.stack([b["pos"] for b in bodies]),
        body_quat=np.stack([b["quat"] for b in bodies]), body_axis=np.stack([b["axis"] for b in bodies]),
        geom_body=np.array([g[0] for g in geoms]), geom_mesh=np.array([g[1] for g in geoms]), hull_adr=np.array(hull_adr),
        hull_pts=np.concatenate(hull_pts), exclude=np.array([(names.index(a), names.index(b)) for a, b in excl]),
        cube_half=vec(cube.find("geom").get("size"), 3), body_names=np.array(names),
        ncube=len(cubes), box_half=np.stack([vec(b.find("geom").get("size"), 3) for b in cubes] + [vec(g.get("size"), 3) for g in walls]),
        wall_pos=np.stack([vec(g.get("pos"), 3) for g in walls]) if walls else np.zeros((0, 3)),
        cube_masses=np.array([float(b.find("inertial").get("mass")) for b in cubes]),
        cube_diaginertias=np.stack([vec(b.find("inertial").get("diaginertia"), 3) for b in cubes]),
        cube_pos0s=np.stack([vec(b.get("pos"), 3) for b in cubes]),
        inertial=np.stack(inertial), joints=np.stack(joints), geom_par=np.array(geom_par), cube_mass=float(cine.get("mass")),
        cube_diaginertia=vec(cine.get("diaginertia"), 3), cube_pos0=vec(cube.get("pos"), 3), act_kp=float(act["kp"]), act_kv=float(act["kv"]),
        timestep=float(opt["timestep"]), impratio=float(opt["impratio"]), site_pos=vec(arm.find("worldbody").find(".//site").get("pos"), 3))
    print("wrote", out, "bodies", len(bodies), "mesh geoms", len(geoms), "hull vertices", hull_adr[-1])


if __name__ == "__main__":
    for t in SCENES:
        main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/gym_lowcostrobot/assets/low_cost_robot_6dof", t)
