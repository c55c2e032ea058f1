"""MJCF loader for the bin-a-cube scene (SO101 arm + gripper, free cube, table mesh, bin).

Parses the same three XML files the reference hands to MuJoCo at
``gym_so100/env.py:111-112`` (``so100_transfer_cube.xml`` which ``<include>``s
``scene_so100.xml`` and ``trs_so_arm100/so_arm100.xml``), resolves ``<default>``
classes / ``childclass``, reads the binary STL collision meshes, builds their
convex hulls and returns a plain-Python :class:`Scene` description.  The
compile-time constants MuJoCo derives in ``mj_setConst`` are computed in
``model.py``; nothing here is on the step path.

Only the MJCF subset these files use is implemented (see SURVEY.md Appendix A,
"MJCF semantics the loader must get right"); unknown physics-relevant
constructs raise instead of being silently ignored.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# MuJoCo geom type enum values (mjtGeom); only box and mesh occur in this scene.
GEOM_BOX = 6
GEOM_MESH = 7
_GEOM_TYPES = {"plane": 0, "hfield": 1, "sphere": 2, "capsule": 3, "ellipsoid": 4,
               "cylinder": 5, "box": 6, "mesh": 7}

JNT_FREE = 0
JNT_HINGE = 3


# --------------------------------------------------------------------------- math
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


def axis_angle_quat(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    s = np.sin(0.5 * angle)
    return np.array([np.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def euler_to_quat(e, seq="xyz"):
    """MuJoCo ``eulerseq``: lower-case letters are intrinsic (rotating-frame) axes."""
    q = np.array([1.0, 0.0, 0.0, 0.0])
    for ch, ang in zip(seq, e):
        ax = {"x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1)}[ch.lower()]
        r = axis_angle_quat(ax, ang)
        q = quat_mul(q, r) if ch.islower() else quat_mul(r, q)
    return q


def _floats(s: str) -> np.ndarray:
    return np.array([float(t) for t in s.split()], dtype=np.float64)


# --------------------------------------------------------------------------- data
@dataclass
class Mesh:
    name: str
    file: str
    scale: np.ndarray
    verts: Optional[np.ndarray] = None       # unique vertices (scaled), authoring frame
    hull: Optional[np.ndarray] = None        # convex-hull vertices (subset of verts)
    hull_faces: Optional[np.ndarray] = None  # triangles into ``hull``


@dataclass
class Geom:
    id: int
    name: str
    body: int
    type: int
    pos: np.ndarray
    quat: np.ndarray
    size: np.ndarray
    contype: int
    conaffinity: int
    condim: int
    friction: np.ndarray
    solref: np.ndarray
    solimp: np.ndarray
    margin: float
    gap: float
    solmix: float
    priority: int
    mesh: Optional[str] = None


@dataclass
class Joint:
    id: int
    name: str
    body: int
    type: int
    axis: np.ndarray
    pos: np.ndarray
    range: np.ndarray
    limited: bool
    frictionloss: float
    armature: float
    damping: float
    stiffness: float
    ref: float
    qposadr: int = 0
    dofadr: int = 0


@dataclass
class Body:
    id: int
    name: str
    parent: int
    pos: np.ndarray
    quat: np.ndarray
    ipos: np.ndarray
    iquat: np.ndarray
    mass: float
    inertia: np.ndarray
    joints: List[int] = field(default_factory=list)
    has_inertial: bool = False


@dataclass
class Site:
    id: int
    name: str
    body: int
    pos: np.ndarray


@dataclass
class Actuator:
    id: int
    name: str
    joint: str
    kp: float
    kv: float            # explicit kv (0 when dampratio is used)
    dampratio: float
    ctrlrange: np.ndarray
    ctrllimited: bool
    forcerange: np.ndarray
    forcelimited: bool
    gear: float


@dataclass
class Scene:
    bodies: List[Body]
    joints: List[Joint]
    geoms: List[Geom]
    sites: List[Site]
    actuators: List[Actuator]
    meshes: Dict[str, Mesh]
    excludes: List[Tuple[str, str]]
    option: Dict[str, object]
    keyframes: Dict[str, Dict[str, np.ndarray]]


# --------------------------------------------------------------------------- STL + hull
def read_binary_stl(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        buf = f.read()
    (ntri,) = struct.unpack("<I", buf[80:84])
    if 84 + 50 * ntri != len(buf):
        raise ValueError(f"{path}: not a binary STL (size {len(buf)} vs {ntri} triangles)")
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    tri = np.frombuffer(buf, dtype=rec, count=ntri, offset=84)
    return tri["v"].reshape(-1, 3).astype(np.float64)


def convex_hull(verts: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Hull vertices (in input order) and outward-oriented triangles (MuJoCo uses qhull too)."""
    from scipy.spatial import ConvexHull

    hull = ConvexHull(verts)
    idx = np.sort(np.unique(hull.simplices))
    remap = -np.ones(len(verts), dtype=np.int64)
    remap[idx] = np.arange(len(idx))
    faces = remap[hull.simplices]
    hv = verts[idx]
    c = hv.mean(axis=0)
    for k in range(len(faces)):
        a, b, cc = hv[faces[k]]
        if np.dot(np.cross(b - a, cc - a), a - c) < 0:
            faces[k] = faces[k][[0, 2, 1]]
    return hv, faces


# --------------------------------------------------------------------------- defaults
class _Defaults:
    """Nested ``<default class=...>`` tree: class -> element tag -> attribute dict."""

    def __init__(self):
        self.classes: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}
        self.parent: Dict[str, Optional[str]] = {"main": None}

    def add_tree(self, elem: ET.Element, parent: Optional[str]):
        name = elem.get("class", "main" if parent is None else None)
        if name is None:
            raise ValueError("nested <default> without class")
        if name not in self.classes:
            self.classes[name] = {}
            self.parent[name] = parent
        if parent is not None and name != "main":
            # inherit a copy of the parent's attributes
            for tag, attrs in self.classes[parent].items():
                merged = dict(attrs)
                merged.update(self.classes[name].get(tag, {}))
                self.classes[name][tag] = merged
        for child in elem:
            if child.tag == "default":
                continue
            d = self.classes[name].setdefault(child.tag, {})
            d.update(child.attrib)
        for child in elem:
            if child.tag == "default":
                self.add_tree(child, name)

    def resolve(self, tag: str, elem: ET.Element, childclass: Optional[str]) -> Dict[str, str]:
        cls = elem.get("class", childclass or "main")
        if cls not in self.classes:
            raise ValueError(f"unknown default class {cls!r}")
        out = dict(self.classes[cls].get(tag, {}))
        out.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return out


# --------------------------------------------------------------------------- parsing
def _expand_includes(path: str) -> Tuple[ET.Element, List[Tuple[ET.Element, str]]]:
    """Return the root with every ``<include>`` spliced in, plus (section, source-dir) pairs."""
    root = ET.parse(path).getroot()
    base = os.path.dirname(os.path.abspath(path))
    sections: List[Tuple[ET.Element, str]] = []

    def walk(elem: ET.Element, src_dir: str):
        for child in list(elem):
            if child.tag == "include":
                inc_path = os.path.join(src_dir, child.get("file"))
                inc_root = ET.parse(inc_path).getroot()
                if inc_root.tag not in ("mujoco", "mujocoinclude"):
                    raise ValueError(f"{inc_path}: unexpected root <{inc_root.tag}>")
                walk(inc_root, os.path.dirname(os.path.abspath(inc_path)))
            else:
                sections.append((child, src_dir))

    walk(root, base)
    return root, sections


def _find_mesh_file(fname: str, src_dir: str, main_dir: str, meshdir: str) -> str:
    # SURVEY 8a-M "Mesh file resolution": meshdir="assets/" does not exist in the checkout,
    # exactly one of these candidates does.
    cands = [os.path.join(src_dir, meshdir, fname), os.path.join(src_dir, fname),
             os.path.join(main_dir, meshdir, fname), os.path.join(main_dir, fname)]
    for c in cands:
        if os.path.isfile(c):
            return c
    raise FileNotFoundError(f"mesh file {fname!r} not found in {cands}")


def load_scene(xml_path: str) -> Scene:
    root, sections = _expand_includes(xml_path)
    main_dir = os.path.dirname(os.path.abspath(xml_path))

    compiler = {"angle": "degree", "meshdir": "", "eulerseq": "xyz", "autolimits": "true"}
    option: Dict[str, object] = {
        "timestep": 0.002, "gravity": np.array([0.0, 0.0, -9.81]), "cone": "pyramidal",
        "impratio": 1.0, "integrator": "Euler", "solver": "Newton", "iterations": 100,
        "tolerance": 1e-8, "ls_iterations": 50, "ls_tolerance": 0.01,
    }
    defaults = _Defaults()
    meshes: Dict[str, Mesh] = {}
    excludes: List[Tuple[str, str]] = []
    keyframes: Dict[str, Dict[str, np.ndarray]] = {}

    # pass 1: compiler / option / default / asset (document order matters for none of these here)
    for sec, src_dir in sections:
        if sec.tag == "compiler":
            compiler.update(sec.attrib)
        elif sec.tag == "option":
            for k, v in sec.attrib.items():
                if k in ("timestep", "impratio", "tolerance", "ls_tolerance"):
                    option[k] = float(v)
                elif k in ("iterations", "ls_iterations"):
                    option[k] = int(v)
                elif k == "gravity":
                    option[k] = _floats(v)
                else:
                    option[k] = v
        elif sec.tag == "default":
            defaults.add_tree(sec, None)
    for sec, src_dir in sections:
        if sec.tag == "asset":
            for m in sec:
                if m.tag != "mesh":
                    continue
                attrs = defaults.resolve("mesh", m, None)
                fname = attrs["file"]
                name = attrs.get("name", os.path.splitext(os.path.basename(fname))[0])
                scale = _floats(attrs.get("scale", "1 1 1"))
                path = _find_mesh_file(fname, src_dir, main_dir, compiler.get("meshdir", ""))
                meshes[name] = Mesh(name=name, file=path, scale=scale)
        elif sec.tag == "contact":
            for e in sec:
                if e.tag == "exclude":
                    excludes.append((e.get("body1"), e.get("body2")))
                else:
                    raise NotImplementedError(f"<contact><{e.tag}>")
        elif sec.tag == "keyframe":
            for k in sec:
                keyframes[k.get("name")] = {a: _floats(v) for a, v in k.attrib.items() if a != "name"}
        elif sec.tag in ("equality", "tendon", "sensor"):
            raise NotImplementedError(f"<{sec.tag}> is not used by the bin-a-cube scene")

    if compiler["angle"] != "radian":
        raise NotImplementedError("only angle=radian is supported (so_arm100.xml:2)")
    eulerseq = compiler.get("eulerseq", "xyz")

    def frame(attrs) -> Tuple[np.ndarray, np.ndarray]:
        pos = _floats(attrs.get("pos", "0 0 0"))
        if "quat" in attrs:
            q = _floats(attrs["quat"])
            q = q / np.linalg.norm(q)
        elif "euler" in attrs:
            q = euler_to_quat(_floats(attrs["euler"]), eulerseq)
        else:
            for bad in ("axisangle", "xyaxes", "zaxis"):
                if bad in attrs:
                    raise NotImplementedError(bad)
            q = np.array([1.0, 0.0, 0.0, 0.0])
        return pos, q

    bodies: List[Body] = [Body(0, "world", 0, np.zeros(3), np.array([1.0, 0, 0, 0]), np.zeros(3),
                               np.array([1.0, 0, 0, 0]), 0.0, np.zeros(3))]
    joints: List[Joint] = []
    geoms: List[Geom] = []
    sites: List[Site] = []

    def add_geom(e, body_id, childclass):
        a = defaults.resolve("geom", e, childclass)
        gtype = _GEOM_TYPES[a.get("type", "sphere")]
        if gtype not in (GEOM_BOX, GEOM_MESH):
            raise NotImplementedError(f"geom type {a.get('type')} (scene has only box and mesh)")
        pos, quat = frame(a)
        size = np.zeros(3)
        if "size" in a:
            s = _floats(a["size"])
            size[:len(s)] = s
        fr = np.array([1.0, 0.005, 0.0001])
        if "friction" in a:
            f = _floats(a["friction"])
            fr[:len(f)] = f
        solref = np.array([0.02, 1.0])
        if "solref" in a:
            s = _floats(a["solref"])
            solref[:len(s)] = s
        solimp = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
        if "solimp" in a:
            s = _floats(a["solimp"])
            solimp[:len(s)] = s   # partial spec keeps the remaining defaults
        gid = len(geoms)
        geoms.append(Geom(
            id=gid, name=a.get("name", ""), body=body_id, type=gtype, pos=pos, quat=quat, size=size,
            contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
            condim=int(a.get("condim", 3)), friction=fr, solref=solref, solimp=solimp,
            margin=float(a.get("margin", 0)), gap=float(a.get("gap", 0)),
            solmix=float(a.get("solmix", 1)), priority=int(a.get("priority", 0)),
            mesh=a.get("mesh")))

    def add_joint(e, body_id, childclass, free=False):
        a = defaults.resolve("joint", e, childclass)
        jtype = JNT_FREE if (free or a.get("type") == "free") else JNT_HINGE
        if a.get("type", "hinge") not in ("hinge", "free"):
            raise NotImplementedError(f"joint type {a.get('type')}")
        rng = _floats(a["range"]) if "range" in a else np.zeros(2)
        limited = a.get("limited", "auto")
        if limited == "auto":
            lim = "range" in a and compiler.get("autolimits", "true") == "true"
        else:
            lim = limited == "true"
        axis = _floats(a.get("axis", "0 0 1"))
        axis = axis / np.linalg.norm(axis)
        joints.append(Joint(
            id=len(joints), name=a.get("name", ""), body=body_id, type=jtype, axis=axis,
            pos=_floats(a.get("pos", "0 0 0")), range=rng, limited=bool(lim and jtype == JNT_HINGE),
            frictionloss=float(a.get("frictionloss", 0)), armature=float(a.get("armature", 0)),
            damping=float(a.get("damping", 0)), stiffness=float(a.get("stiffness", 0)),
            ref=float(a.get("ref", 0))))
        bodies[body_id].joints.append(joints[-1].id)

    def walk_body(e, parent_id, childclass):
        childclass = e.get("childclass", childclass)
        pos, quat = frame(e.attrib)
        bid = len(bodies)
        if e.get("mocap", "false") == "true":
            raise NotImplementedError("mocap bodies (EE scene) are out of scope")
        body = Body(bid, e.get("name", f"body{bid}"), parent_id, pos, quat, np.zeros(3),
                    np.array([1.0, 0, 0, 0]), 0.0, np.zeros(3))
        bodies.append(body)
        # MuJoCo assigns ids depth-first in document order per element kind
        for c in e:
            if c.tag == "inertial":
                ipos, iquat = frame(c.attrib)
                body.ipos, body.iquat = ipos, iquat
                body.mass = float(c.get("mass"))
                if "diaginertia" in c.attrib:
                    body.inertia = _floats(c.get("diaginertia"))
                else:
                    raise NotImplementedError("fullinertia")
                body.has_inertial = True
            elif c.tag == "joint":
                add_joint(c, bid, childclass)
            elif c.tag == "freejoint":
                add_joint(c, bid, childclass, free=True)
            elif c.tag == "geom":
                add_geom(c, bid, childclass)
            elif c.tag == "site":
                a = defaults.resolve("site", c, childclass)
                sites.append(Site(len(sites), a.get("name", ""), bid, _floats(a.get("pos", "0 0 0"))))
            elif c.tag in ("camera", "light"):
                pass  # rendering only (declared out of scope)
            elif c.tag == "body":
                pass
            else:
                raise NotImplementedError(f"<body><{c.tag}>")
        for c in e:
            if c.tag == "body":
                walk_body(c, bid, childclass)

    # world-level geoms/sites first would get lower ids; this scene has none outside bodies.
    for sec, _ in sections:
        if sec.tag == "worldbody":
            for c in sec:
                if c.tag == "geom":
                    add_geom(c, 0, None)
                elif c.tag == "site":
                    a = defaults.resolve("site", c, None)
                    sites.append(Site(len(sites), a.get("name", ""), 0, _floats(a.get("pos", "0 0 0"))))
    # MuJoCo numbers elements depth-first over the merged worldbody; geoms of a body precede
    # those of its children.  walk_body() adds a body's own elements before recursing.
    for sec, _ in sections:
        if sec.tag == "worldbody":
            for c in sec:
                if c.tag == "body":
                    walk_body(c, 0, None)

    # joint addresses
    nq = nv = 0
    for j in joints:
        j.qposadr, j.dofadr = nq, nv
        nq += 7 if j.type == JNT_FREE else 1
        nv += 6 if j.type == JNT_FREE else 1
    for b in bodies:
        if len(b.joints) > 1:
            raise NotImplementedError("more than one joint per body")

    # actuators
    actuators: List[Actuator] = []
    jname = {j.name: j for j in joints}
    for sec, _ in sections:
        if sec.tag != "actuator":
            continue
        for e in sec:
            if e.tag != "position":
                raise NotImplementedError(f"actuator <{e.tag}>")
            a = defaults.resolve("position", e, None)
            j = jname[a["joint"]]
            kp = float(a.get("kp", 1))
            inherit = float(a.get("inheritrange", 0))
            if "ctrlrange" in a:
                cr = _floats(a["ctrlrange"])
                climited = True
            elif inherit > 0:
                mid, half = 0.5 * (j.range[0] + j.range[1]), 0.5 * (j.range[1] - j.range[0])
                cr = np.array([mid - half * inherit, mid + half * inherit])
                climited = True
            else:
                cr, climited = np.zeros(2), False
            fr = _floats(a["forcerange"]) if "forcerange" in a else np.zeros(2)
            actuators.append(Actuator(
                id=len(actuators), name=a.get("name", ""), joint=j.name, kp=kp,
                kv=float(a.get("kv", 0)), dampratio=float(a.get("dampratio", 0)),
                ctrlrange=cr, ctrllimited=climited, forcerange=fr, forcelimited="forcerange" in a,
                gear=float(a.get("gear", "1").split()[0])))

    # meshes actually used for collision get vertices + hulls
    used = {g.mesh for g in geoms if g.type == GEOM_MESH and (g.contype or g.conaffinity)}
    for name in used:
        m = meshes[name]
        raw = read_binary_stl(m.file) * m.scale[None, :]
        m.verts = np.unique(raw, axis=0)
        m.hull, m.hull_faces = convex_hull(m.verts)

    return Scene(bodies=bodies, joints=joints, geoms=geoms, sites=sites, actuators=actuators,
                 meshes=meshes, excludes=excludes, option=option, keyframes=keyframes)
