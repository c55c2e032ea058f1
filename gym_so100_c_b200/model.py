"""Compile a parsed :class:`mjcf.Scene` into the flat constant model both back ends read.

This is the one-off "model compile" the reference delegates to MuJoCo at
``gym_so100/env.py:98-112`` (``mujoco.Physics.from_xml_path``): qpos0 kinematics,
``dof_M0``, position-actuator ``kv`` from ``dampratio``, ``dof_invweight0`` /
``body_invweight0``, ``meaninertia`` (the ``mj_setConst`` equivalents of SURVEY.md
Appendix A), the static list of candidate collision pairs with their mixed contact
parameters, convex-hull vertex pools and the task constants of
``gym_so100/constants.py`` / ``gym_so100/utils.py`` / ``gym_so100/tasks/single_arm.py``.

The result is packed as one little-endian C struct (``MODEL_DTYPE``); the matching
C declaration is generated into ``include/so100_model.h`` by :func:`c_header` so the
CUDA library, the C oracle and Python cannot drift apart.  All reals are float64 in
the blob; the CUDA library narrows to float32 when it uploads its constant struct.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from . import mjcf
from .mjcf import GEOM_BOX, GEOM_MESH, JNT_FREE, JNT_HINGE, quat_mul, quat_to_mat, axis_angle_quat

MAGIC = 0x53313030  # 'S100'
VERSION = 3

MAXBODY = 16
MAXDOF = 12
MAXQ = 13
MAXACT = 6
MAXGEOM = 32
MAXPAIR = 192
MAXVERT = 2560
MAXSITE = 8

_f8, _i4 = "<f8", "<i4"

MODEL_DTYPE = np.dtype([
    # ---- header
    ("magic", "<u4"), ("version", "<u4"),
    ("nbody", _i4), ("nq", _i4), ("nv", _i4), ("nu", _i4), ("ngeom", _i4), ("ngeom_all", _i4),
    ("nsite", _i4), ("npair", _i4), ("nvert", _i4), ("nsubstep", _i4),
    ("iterations", _i4), ("ls_iterations", _i4),
    # ---- options (so_arm100.xml:4 + MuJoCo defaults)
    ("timestep", _f8), ("gravity", _f8, 3), ("impratio", _f8), ("tolerance", _f8),
    ("ls_tolerance", _f8), ("meaninertia", _f8),
    # ---- bodies
    ("body_parent", _i4, MAXBODY), ("body_jtype", _i4, MAXBODY), ("body_dofadr", _i4, MAXBODY),
    ("body_qposadr", _i4, MAXBODY), ("body_weldid", _i4, MAXBODY),
    ("body_pos", _f8, (MAXBODY, 3)), ("body_quat", _f8, (MAXBODY, 4)),
    ("body_ipos", _f8, (MAXBODY, 3)), ("body_iquat", _f8, (MAXBODY, 4)),
    ("body_mass", _f8, MAXBODY), ("body_inertia", _f8, (MAXBODY, 3)),
    ("body_jaxis", _f8, (MAXBODY, 3)), ("body_invweight0", _f8, (MAXBODY, 2)),
    # ---- dofs
    ("dof_body", _i4, MAXDOF), ("dof_limited", _i4, MAXDOF),
    ("dof_armature", _f8, MAXDOF), ("dof_frictionloss", _f8, MAXDOF),
    ("dof_invweight0", _f8, MAXDOF), ("dof_M0", _f8, MAXDOF), ("dof_range", _f8, (MAXDOF, 2)),
    ("qpos0", _f8, MAXQ),
    # ---- position actuators
    ("act_dof", _i4, MAXACT), ("act_kp", _f8, MAXACT), ("act_kv", _f8, MAXACT),
    ("act_ctrlrange", _f8, (MAXACT, 2)), ("act_forcerange", _f8, (MAXACT, 2)),
    # ---- collidable geoms (index "cg"; geom_mjid is the MuJoCo geom id)
    ("geom_mjid", _i4, MAXGEOM), ("geom_body", _i4, MAXGEOM), ("geom_type", _i4, MAXGEOM),
    ("geom_vadr", _i4, MAXGEOM), ("geom_vnum", _i4, MAXGEOM),
    ("geom_pos", _f8, (MAXGEOM, 3)), ("geom_quat", _f8, (MAXGEOM, 4)), ("geom_size", _f8, (MAXGEOM, 3)),
    ("geom_center", _f8, (MAXGEOM, 3)), ("geom_half", _f8, (MAXGEOM, 3)), ("geom_rbound", _f8, MAXGEOM),
    # ---- hull vertex pool (body frame)
    ("vert", _f8, (MAXVERT, 3)),
    # ---- candidate geom pairs, geom1 has the lower (type, id)
    ("pair_g1", _i4, MAXPAIR), ("pair_g2", _i4, MAXPAIR), ("pair_condim", _i4, MAXPAIR),
    ("pair_friction", _f8, (MAXPAIR, 3)), ("pair_solref", _f8, (MAXPAIR, 2)),
    ("pair_solimp", _f8, (MAXPAIR, 5)),
    # ---- sites
    ("site_body", _i4, MAXSITE), ("site_pos", _f8, (MAXSITE, 3)),
    # ---- task constants
    ("site_cube", _i4), ("site_ee", _i4), ("site_bin", _i4),
    ("cg_cube", _i4), ("cg_table", _i4), ("pad_mask", "<u4"),
    ("max_episode_steps", _i4), ("goal_curriculum_steps", _i4),
    ("start_pose", _f8, 6), ("act_lo", _f8, 6), ("act_hi", _f8, 6),
    ("box_lo", _f8, 3), ("box_hi", _f8, 3),
    ("bin_hw", _f8), ("bin_h", _f8), ("cube_half", _f8), ("goal_threshold", _f8),
    ("bin_goal_lo", _f8, 3), ("bin_goal_hi", _f8, 3),
    ("lift_goal_xy", _f8), ("lift_goal_zlo", _f8), ("lift_goal_zhi", _f8),
], align=True)


# --------------------------------------------------------------------------- task constants
# gym_so100/constants.py:32-39
SO100_START_ARM_POSE = (0.0, -0.96, 1.16, 0.0, 0.0, 0.02239)
# gym_so100/constants.py:78-86 (ranges handed to `unnormalize`)
ACTION_RANGES = ((-1.92, 1.92), (-3.32, 0.174), (-0.174, 3.14), (-1.66, 1.66), (-2.79, 2.79), (-0.174, 1.75))
# gym_so100/utils.py:18-21
BOX_RANGE_LO = (-0.25, 0.3, 0.05)
BOX_RANGE_HI = (-0.15, 0.6, 0.05)
# gym_so100/constants.py:29-30 (float32 arrays in the reference)
BIN_MIN = np.array([-0.25, 0.7, 0.01], dtype=np.float32)
BIN_MAX = np.array([-0.14, 0.76, 0.05], dtype=np.float32)
DT = 0.02  # gym_so100/constants.py:4


# --------------------------------------------------------------------------- numpy kinematics
def fk(m: np.ndarray, qpos: np.ndarray):
    """Body frames (xpos, xquat) for ``qpos``; used for compile-time constants and tests only."""
    nb = int(m["nbody"])
    xpos = np.zeros((nb, 3))
    xquat = np.zeros((nb, 4))
    xquat[0] = (1, 0, 0, 0)
    for b in range(1, nb):
        p = int(m["body_parent"][b])
        jt = int(m["body_jtype"][b])
        if jt == JNT_FREE:
            a = int(m["body_qposadr"][b])
            xpos[b] = qpos[a:a + 3]
            q = qpos[a + 3:a + 7]
            xquat[b] = q / np.linalg.norm(q)
            continue
        R = quat_to_mat(xquat[p])
        xpos[b] = xpos[p] + R @ m["body_pos"][b]
        xquat[b] = quat_mul(xquat[p], m["body_quat"][b])
        if jt == JNT_HINGE:
            a = int(m["body_qposadr"][b])
            ang = qpos[a] - m["qpos0"][a]
            xquat[b] = quat_mul(xquat[b], axis_angle_quat(m["body_jaxis"][b], ang))
    return xpos, xquat


def jacobians(m: np.ndarray, xpos, xquat, body: int, point: np.ndarray):
    """(jacp, jacr) 3 x nv of a world ``point`` attached to ``body``."""
    nv = int(m["nv"])
    jp = np.zeros((3, nv))
    jr = np.zeros((3, nv))
    b = body
    while b > 0:
        jt = int(m["body_jtype"][b])
        d = int(m["body_dofadr"][b])
        R = quat_to_mat(xquat[b])
        if jt == JNT_HINGE:
            ax = R @ m["body_jaxis"][b]
            jr[:, d] = ax
            jp[:, d] = np.cross(ax, point - xpos[b])
        elif jt == JNT_FREE:
            jp[:, d:d + 3] = np.eye(3)
            for k in range(3):
                jr[:, d + 3 + k] = R[:, k]
                jp[:, d + 3 + k] = np.cross(R[:, k], point - xpos[b])
        b = int(m["body_parent"][b])
    return jp, jr


def mass_matrix(m: np.ndarray, qpos: np.ndarray) -> np.ndarray:
    """M(q) = sum_b J_b^T I_b J_b + armature (definition, not CRB)."""
    nb, nv = int(m["nbody"]), int(m["nv"])
    xpos, xquat = fk(m, qpos)
    M = np.zeros((nv, nv))
    for b in range(1, nb):
        mass = float(m["body_mass"][b])
        if mass <= 0:
            continue
        R = quat_to_mat(xquat[b])
        com = xpos[b] + R @ m["body_ipos"][b]
        Ri = quat_to_mat(quat_mul(xquat[b], m["body_iquat"][b]))
        Iw = Ri @ np.diag(m["body_inertia"][b]) @ Ri.T
        jp, jr = jacobians(m, xpos, xquat, b, com)
        M += mass * jp.T @ jp + jr.T @ Iw @ jr
    M[np.arange(nv), np.arange(nv)] += m["dof_armature"][:nv]
    return M


def site_xpos(m: np.ndarray, qpos: np.ndarray) -> np.ndarray:
    xpos, xquat = fk(m, qpos)
    out = np.zeros((int(m["nsite"]), 3))
    for s in range(int(m["nsite"])):
        b = int(m["site_body"][s])
        out[s] = xpos[b] + quat_to_mat(xquat[b]) @ m["site_pos"][s]
    return out


# --------------------------------------------------------------------------- compile
def _mix_pair(g1: mjcf.Geom, g2: mjcf.Geom):
    """MuJoCo contact-parameter mixing for equal priority (SURVEY 8a-M "Pair mixing")."""
    if g1.priority != g2.priority:
        raise NotImplementedError("geom priority")
    condim = max(g1.condim, g2.condim)
    friction = np.maximum(g1.friction, g2.friction)
    s1, s2 = g1.solmix, g2.solmix
    mix = s1 / (s1 + s2) if (s1 + s2) > 0 else 0.5
    if g1.solref[0] > 0 and g2.solref[0] > 0:
        solref = mix * g1.solref + (1 - mix) * g2.solref
    else:
        solref = np.minimum(g1.solref, g2.solref)
    solimp = mix * g1.solimp + (1 - mix) * g2.solimp
    return condim, friction, solref, solimp


def compile_model(scene: mjcf.Scene) -> np.ndarray:
    m = np.zeros((), dtype=MODEL_DTYPE)
    opt = scene.option
    if opt["cone"] != "elliptic" or opt["integrator"] != "Euler" or opt["solver"] != "Newton":
        raise NotImplementedError("kernels implement cone=elliptic, integrator=Euler, solver=Newton")
    nb = len(scene.bodies)
    nq = sum(7 if j.type == JNT_FREE else 1 for j in scene.joints)
    nv = sum(6 if j.type == JNT_FREE else 1 for j in scene.joints)
    nu = len(scene.actuators)
    if nb > MAXBODY or nv > MAXDOF or nq > MAXQ or nu > MAXACT or len(scene.sites) > MAXSITE:
        raise ValueError("scene exceeds the compiled capacities")
    m["magic"], m["version"] = MAGIC, VERSION
    m["nbody"], m["nq"], m["nv"], m["nu"] = nb, nq, nv, nu
    m["timestep"], m["gravity"], m["impratio"] = opt["timestep"], opt["gravity"], opt["impratio"]
    m["tolerance"], m["ls_tolerance"] = opt["tolerance"], opt["ls_tolerance"]
    m["iterations"], m["ls_iterations"] = opt["iterations"], opt["ls_iterations"]
    # dm_control: n_sub_steps = round(control_timestep / timestep)  (env.py:120-127)
    m["nsubstep"] = int(round(DT / opt["timestep"]))

    m["body_jtype"][:] = -1
    m["body_dofadr"][:] = -1
    m["body_qposadr"][:] = -1
    m["body_quat"][:, 0] = 1
    m["body_iquat"][:, 0] = 1
    weld = [0] * nb
    for b in scene.bodies:
        i = b.id
        m["body_parent"][i] = b.parent
        m["body_pos"][i], m["body_quat"][i] = b.pos, b.quat
        m["body_ipos"][i], m["body_iquat"][i] = b.ipos, b.iquat
        m["body_mass"][i], m["body_inertia"][i] = b.mass, b.inertia
        if b.joints:
            j = scene.joints[b.joints[0]]
            if np.linalg.norm(j.pos) != 0 or j.ref != 0 or j.damping != 0 or j.stiffness != 0:
                raise NotImplementedError("joint pos/ref/damping/stiffness (all zero in this scene)")
            m["body_jtype"][i] = j.type
            m["body_dofadr"][i], m["body_qposadr"][i] = j.dofadr, j.qposadr
            m["body_jaxis"][i] = j.axis
            weld[i] = i
            if b.mass <= 0:
                raise ValueError(f"moving body {b.name} needs an explicit <inertial>")
            nd = 6 if j.type == JNT_FREE else 1
            for k in range(nd):
                d = j.dofadr + k
                m["dof_body"][d] = i
                m["dof_armature"][d] = j.armature
                m["dof_frictionloss"][d] = j.frictionloss
                m["dof_limited"][d] = int(j.limited)
                m["dof_range"][d] = j.range
        else:
            weld[i] = weld[b.parent] if i > 0 else 0
        m["body_weldid"][i] = weld[i]

    # qpos0: hinges at ref (0), free joint at the body's MJCF pose
    for j in scene.joints:
        if j.type == JNT_FREE:
            b = scene.bodies[j.body]
            if b.parent != 0:
                raise NotImplementedError("free joint below a non-world body")
            m["qpos0"][j.qposadr:j.qposadr + 3] = b.pos
            m["qpos0"][j.qposadr + 3:j.qposadr + 7] = b.quat

    # sites
    m["nsite"] = len(scene.sites)
    for s in scene.sites:
        m["site_body"][s.id], m["site_pos"][s.id] = s.body, s.pos
    names = {s.name: s.id for s in scene.sites}
    m["site_cube"], m["site_ee"], m["site_bin"] = names["cube_site"], names["ee_site"], names["bin_center"]

    # mj_setConst equivalents at qpos0
    qpos0 = m["qpos0"][:nq].copy()
    M0 = mass_matrix(m, qpos0)
    Minv = np.linalg.inv(M0)
    m["dof_M0"][:nv] = np.diag(M0)
    m["meaninertia"] = float(np.mean(np.diag(M0)))
    inv = np.diag(Minv).copy()
    for j in scene.joints:
        if j.type == JNT_FREE:
            d = j.dofadr
            inv[d:d + 3] = inv[d:d + 3].mean()
            inv[d + 3:d + 6] = inv[d + 3:d + 6].mean()
    m["dof_invweight0"][:nv] = inv
    xpos, xquat = fk(m, qpos0)
    for b in range(1, nb):
        if weld[b] == 0:
            continue
        com = xpos[b] + quat_to_mat(xquat[b]) @ m["body_ipos"][b]
        jp, jr = jacobians(m, xpos, xquat, b, com)
        m["body_invweight0"][b, 0] = np.trace(jp @ Minv @ jp.T) / 3
        m["body_invweight0"][b, 1] = np.trace(jr @ Minv @ jr.T) / 3

    # position actuators: kv = dampratio * 2 * sqrt(kp * reflected inertia)
    jname = {j.name: j for j in scene.joints}
    for a in scene.actuators:
        j = jname[a.joint]
        if j.type != JNT_HINGE or a.gear != 1 or not a.ctrllimited or not a.forcelimited:
            raise NotImplementedError("actuator form (scene uses hinge/gear 1/ctrl+force limited)")
        kv = a.kv
        if a.dampratio > 0:
            kv = a.dampratio * 2.0 * np.sqrt(a.kp * M0[j.dofadr, j.dofadr])
        m["act_dof"][a.id] = j.dofadr
        m["act_kp"][a.id], m["act_kv"][a.id] = a.kp, kv
        m["act_ctrlrange"][a.id], m["act_forcerange"][a.id] = a.ctrlrange, a.forcerange

    # collidable geoms + hull vertex pools (vertices baked into the body frame)
    coll = [g for g in scene.geoms if g.contype or g.conaffinity]
    if len(coll) > MAXGEOM:
        raise ValueError("too many collidable geoms")
    m["ngeom"], m["ngeom_all"] = len(coll), len(scene.geoms)
    m["geom_quat"][:, 0] = 1
    vadr = 0
    cg_of = {}
    for cg, g in enumerate(coll):
        cg_of[g.id] = cg
        m["geom_mjid"][cg], m["geom_body"][cg], m["geom_type"][cg] = g.id, g.body, g.type
        if g.margin != 0 or g.gap != 0:
            raise NotImplementedError("geom margin/gap")
        if g.type == GEOM_BOX:
            m["geom_pos"][cg], m["geom_quat"][cg], m["geom_size"][cg] = g.pos, g.quat, g.size
            m["geom_center"][cg], m["geom_half"][cg] = g.pos, g.size
            m["geom_rbound"][cg] = np.linalg.norm(g.size)
            m["geom_vadr"][cg], m["geom_vnum"][cg] = -1, 0
        else:
            hv = scene.meshes[g.mesh].hull
            v = g.pos[None, :] + hv @ quat_to_mat(g.quat).T
            n = len(v)
            if vadr + n > MAXVERT:
                raise ValueError("hull vertex pool overflow")
            m["vert"][vadr:vadr + n] = v
            m["geom_vadr"][cg], m["geom_vnum"][cg] = vadr, n
            lo, hi = v.min(axis=0), v.max(axis=0)
            c = 0.5 * (lo + hi)
            m["geom_center"][cg], m["geom_half"][cg] = c, 0.5 * (hi - lo)
            m["geom_pos"][cg] = c          # mesh geoms: frame = body axes at the AABB centre
            m["geom_rbound"][cg] = np.sqrt(((v - c) ** 2).sum(axis=1).max())
            vadr += n
    m["nvert"] = vadr

    # candidate pairs (SURVEY 8a-M "Collision filtering")
    bname = {b.name: b.id for b in scene.bodies}
    excl = {tuple(sorted((bname[a], bname[b]))) for a, b in scene.excludes}
    pairs = []
    for i, ga in enumerate(coll):
        for gb in coll[i + 1:]:
            b1, b2 = ga.body, gb.body
            if b1 == b2:
                continue
            if not ((ga.contype & gb.conaffinity) or (gb.contype & ga.conaffinity)):
                continue
            w1, w2 = weld[b1], weld[b2]
            if w1 == w2:
                continue  # same weld group (covers static-static: both 0)
            if w1 != 0 and w2 != 0:
                pw1 = weld[scene.bodies[w1].parent]
                pw2 = weld[scene.bodies[w2].parent]
                if pw1 == w2 or pw2 == w1:
                    continue  # parent-child filter
            if tuple(sorted((b1, b2))) in excl:
                continue
            g1, g2 = (ga, gb) if (ga.type, ga.id) <= (gb.type, gb.id) else (gb, ga)
            pairs.append((g1, g2))
    if len(pairs) > MAXPAIR:
        raise ValueError("too many candidate pairs")
    m["npair"] = len(pairs)
    for k, (g1, g2) in enumerate(pairs):
        condim, fr, solref, solimp = _mix_pair(g1, g2)
        if condim not in (3, 4):
            raise NotImplementedError("condim other than 3/4")
        m["pair_g1"][k], m["pair_g2"][k], m["pair_condim"][k] = cg_of[g1.id], cg_of[g2.id], condim
        m["pair_friction"][k], m["pair_solref"][k], m["pair_solimp"][k] = fr, solref, solimp

    # task constants
    gname = {g.name: cg_of[g.id] for g in coll if g.name}
    m["cg_cube"], m["cg_table"] = gname["red_box"], gname["table"]
    pad = 0
    for side in ("fixed", "moving"):
        for i in range(1, 5):   # single_arm.py:334-336
            pad |= 1 << gname[f"{side}_jaw_pad_{i}"]
    m["pad_mask"] = pad
    m["max_episode_steps"] = 700          # gym_so100/__init__.py:27 (CubeToBin); GoalEnv uses 300
    m["goal_curriculum_steps"] = 5000     # gym_so100/env.py:324
    m["start_pose"] = SO100_START_ARM_POSE
    m["act_lo"] = [r[0] for r in ACTION_RANGES]
    m["act_hi"] = [r[1] for r in ACTION_RANGES]
    m["box_lo"], m["box_hi"] = BOX_RANGE_LO, BOX_RANGE_HI
    m["bin_hw"], m["bin_h"], m["cube_half"] = 0.06, 0.03, 0.01      # single_arm.py:68-75
    m["goal_threshold"] = 0.01                                      # env.py:252
    # env.py:245-249: float32 Box built from float32 bin_min/bin_max +- 0.005
    lo = np.array([BIN_MIN[0] + 0.005, BIN_MIN[1] + 0.005, 0.01]).astype(np.float32)
    hi = np.array([BIN_MAX[0] - 0.005, BIN_MAX[1] - 0.005, 0.05]).astype(np.float32)
    m["bin_goal_lo"], m["bin_goal_hi"] = lo.astype(np.float64), hi.astype(np.float64)
    m["lift_goal_xy"], m["lift_goal_zlo"], m["lift_goal_zhi"] = 0.03, 0.01, 0.05   # env.py:325-329
    return m


# --------------------------------------------------------------------------- io
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "bin_a_cube.model")


def pack(m: np.ndarray) -> bytes:
    return np.ascontiguousarray(m).tobytes()


def unpack(buf: bytes) -> np.ndarray:
    if len(buf) != MODEL_DTYPE.itemsize:
        raise ValueError(f"model blob is {len(buf)} bytes, expected {MODEL_DTYPE.itemsize}")
    m = np.frombuffer(buf, dtype=MODEL_DTYPE, count=1)[0].copy()
    if int(m["magic"]) != MAGIC or int(m["version"]) != VERSION:
        raise ValueError("model blob magic/version mismatch")
    return m


def load_model(assets_dir: str | None = None) -> np.ndarray:
    """Compile from a reference ``gym_so100/assets`` directory, or read the committed blob."""
    if assets_dir is not None:
        return compile_model(mjcf.load_scene(os.path.join(assets_dir, "so100_transfer_cube.xml")))
    with open(_DATA, "rb") as f:
        return unpack(f.read())


def c_header() -> str:
    """C declaration of MODEL_DTYPE (written to include/so100_model.h by tools/build_model.py)."""
    lines = [
        "/* GENERATED by gym_so100_c_b200/model.py:c_header() -- do not edit.",
        " * Flat constant model of the bin-a-cube scene (the one-off MuJoCo model compile of",
        " * gym_so100/env.py:98-112).  Little-endian, natural alignment, float64 reals. */",
        "#ifndef SO100_MODEL_H_", "#define SO100_MODEL_H_", "#include <stdint.h>", "",
        f"#define SO100_MODEL_MAGIC 0x{MAGIC:08x}u", f"#define SO100_MODEL_VERSION {VERSION}",
        f"#define SO100_MAXBODY {MAXBODY}", f"#define SO100_MAXDOF {MAXDOF}", f"#define SO100_MAXQ {MAXQ}",
        f"#define SO100_MAXACT {MAXACT}", f"#define SO100_MAXGEOM {MAXGEOM}", f"#define SO100_MAXPAIR {MAXPAIR}",
        f"#define SO100_MAXVERT {MAXVERT}", f"#define SO100_MAXSITE {MAXSITE}",
        "#define SO100_GEOM_BOX 6", "#define SO100_GEOM_MESH 7",
        "#define SO100_JNT_FREE 0", "#define SO100_JNT_HINGE 3", "",
        "typedef struct so100_model {",
    ]
    ctype = {"<f8": "double", "<i4": "int32_t", "<u4": "uint32_t"}
    for name in MODEL_DTYPE.names:
        dt, off = MODEL_DTYPE.fields[name][:2]
        base = dt.base.str
        dims = "".join(f"[{d}]" for d in dt.shape)
        lines.append(f"  {ctype[base]} {name}{dims};  /* offset {off} */")
    lines += ["} so100_model;", "", f"#define SO100_MODEL_BYTES {MODEL_DTYPE.itemsize}",
              "#endif  /* SO100_MODEL_H_ */", ""]
    return "\n".join(lines)
