"""Scene tables of the low-resolution renderer (``so100_render_config``) built from the packed model, and a plain numpy
restatement of the same ray-caster (used by the tests and to document exactly what the kernel draws).

The reference's pixel observation is ``physics.render(height, width, camera_id="top")`` (gym_so100/env.py:130-136,
gym_so100/tasks/single_arm.py:87-91).  Cameras and lights below restate gym_so100/assets/scene_so100.xml:6-30: ``targetbody``
cameras aimed at the table body (0, 0.6, 0), fovy 78 degrees, headlight ambient 0.4 (diffuse: MuJoCo's default 0.4), three
directional lights of diffuse 0.3.  The arm is drawn by its collision hulls (convex hulls of the same STL parts MuJoCo shows as
visual meshes), the finger pads -- collision-only boxes of group 3, invisible in MuJoCo -- are skipped.
"""
from __future__ import annotations

import numpy as np

from .mjcf import quat_to_mat

TABLE_POS = np.array([0.0, 0.6, 0.0])                    # scene_so100.xml:19
CAMERAS = {                                              # scene_so100.xml:26-29 (pos, target, fovy); front_close follows the gripper: not offered
    "top": (np.array([0.0, 0.6, 0.8]), TABLE_POS, 78.0),
    "angle": (np.array([0.0, 0.0, 0.6]), TABLE_POS, 78.0),
    "left_pillar": (np.array([-0.5, 0.2, 0.6]), TABLE_POS, 78.0),
    "right_pillar": (np.array([0.5, 0.2, 0.6]), TABLE_POS, 78.0),
}
LIGHTS = np.array([[1, 1, -1, 0.3], [-1, 1, -1, 0.3], [0, -1, -1, 0.3]], dtype=np.float32)   # scene_so100.xml:13-17: dir, diffuse
HEAD_AMBIENT, HEAD_DIFFUSE = 0.4, 0.4                    # scene_so100.xml:9; MuJoCo default headlight diffuse
RGB_TABLE, RGB_CUBE, RGB_STATIC_BOX, RGB_ARM = (0.2, 0.2, 0.2), (1.0, 0.0, 0.0), (0.5, 0.5, 0.5), (1.0, 1.0, 1.0)


def camera_frame(pos, target):
    """MuJoCo ``targetbody`` camera: z points from the target to the camera, x = (0,0,1) x z (the unit x axis when that
    vanishes, i.e. for a camera straight above its target), y = z x x.  Returns (pos, x, y, z)."""
    z = np.asarray(pos, float) - np.asarray(target, float)
    z /= np.linalg.norm(z)
    x = np.cross([0.0, 0.0, 1.0], z)
    nx = np.linalg.norm(x)
    x = x / nx if nx > 1e-12 else np.array([1.0, 0.0, 0.0])
    y = np.cross(z, x)
    return np.asarray(pos, float), x, y, z


def scene_tables(m: np.ndarray):
    """(planes [P,4], adr [ngeom], num [ngeom], rgb [ngeom,3]) for the collidable geoms in library order.
    num > 0: convex hull drawn from `num` facet planes n.x <= w in the body frame; 0: box; -1: not drawn (finger pads)."""
    from scipy.spatial import ConvexHull
    ng = int(m["ngeom"])
    planes, adr, num, rgb = [], np.zeros(ng, np.int32), np.zeros(ng, np.int32), np.zeros((ng, 3), np.float32)
    pad_mask, cube, table = int(m["pad_mask"]), int(m["cg_cube"]), int(m["cg_table"])
    for g in range(ng):
        nv = int(m["geom_vnum"][g])
        static = int(m["body_weldid"][int(m["geom_body"][g])]) == 0
        rgb[g] = RGB_TABLE if g == table else RGB_CUBE if g == cube else (RGB_STATIC_BOX if (static and nv == 0) else RGB_ARM)
        if (pad_mask >> g) & 1:
            num[g] = -1
            continue
        if nv == 0 or g == table:                       # box geom, or the table top whose hull is an exact cuboid
            num[g] = 0
            continue
        a = int(m["geom_vadr"][g])
        V = np.asarray(m["vert"][a:a + nv], float)
        eq = ConvexHull(V).equations                    # n . x + off <= 0 inside
        keep = []
        for e in eq:                                    # qhull triangulates: merge coplanar facets
            if not any(np.dot(e[:3], q[:3]) > 1 - 1e-10 and abs(e[3] - q[3]) < 1e-9 for q in keep):
                keep.append(e)
        adr[g], num[g] = len(planes), len(keep)
        planes += [[e[0], e[1], e[2], -e[3]] for e in keep]
    return np.asarray(planes, np.float32).reshape(-1, 4), adr, num, rgb


def render_numpy(m: np.ndarray, qpos: np.ndarray, width: int, height: int, camera: str = "top") -> np.ndarray:
    """float64 restatement of csrc/so100_render.cuh for one env: uint8 [height, width, 3]."""
    from . import model as _model
    planes, adr, num, rgb = scene_tables(m)
    planes = planes.astype(np.float64)
    xpos, xquat = _model.fk(m, np.asarray(qpos, float))
    pos, cx, cy, cz = camera_frame(*CAMERAS[camera][:2])
    th = np.tan(0.5 * np.deg2rad(CAMERAS[camera][2]))
    aspect = width / height
    cols, rows = np.meshgrid(np.arange(width), np.arange(height))
    px = th * aspect * (2 * (cols + 0.5) / width - 1)
    py = th * (1 - 2 * (rows + 0.5) / height)
    D = px[..., None] * cx + py[..., None] * cy - cz
    D /= np.linalg.norm(D, axis=-1, keepdims=True)
    best = np.full((height, width), np.inf)
    hit = np.full((height, width), -1)
    nrm = np.zeros((height, width, 3))
    for g in range(int(m["ngeom"])):
        if num[g] < 0:
            continue
        b = int(m["geom_body"][g])
        R = quat_to_mat(xquat[b])
        if num[g] == 0:
            org = xpos[b] + R @ np.asarray(m["geom_center"][g], float)
        else:
            org = xpos[b]
        ol = R.T @ (pos - org)
        dl = D @ R                                       # R^T d for every ray
        tn = np.zeros((height, width)); tf = np.full((height, width), np.inf); nl = np.zeros((height, width, 3))
        miss = np.zeros((height, width), bool)
        if num[g] == 0:
            h = np.asarray(m["geom_half"][g], float)
            for k in range(3):
                dk = dl[..., k]
                par = np.abs(dk) < 1e-12
                miss |= par & (abs(ol[k]) > h[k])
                inv = 1.0 / np.where(par, 1.0, dk)
                t1, t2 = (-h[k] - ol[k]) * inv, (h[k] - ol[k]) * inv
                sgn = np.where(t1 <= t2, -1.0, 1.0)
                lo, hi = np.minimum(t1, t2), np.maximum(t1, t2)
                upd = (~par) & (lo > tn)
                tn = np.where(upd, lo, tn)
                e = np.zeros(3); e[k] = 1.0
                nl = np.where(upd[..., None], sgn[..., None] * e, nl)
                tf = np.where(par, tf, np.minimum(tf, hi))
        else:
            for pl in planes[adr[g]:adr[g] + num[g]]:
                den = dl @ pl[:3]
                dist = pl[3] - ol @ pl[:3]
                ent = den < 0
                tt = dist / np.where(den == 0, 1.0, den)
                upd = ent & (tt > tn)
                tn = np.where(upd, tt, tn)
                nl = np.where(upd[..., None], pl[:3], nl)
                tf = np.where(den > 0, np.minimum(tf, tt), tf)
                miss |= (den == 0) & (dist < 0)
        ok = (~miss) & (tn <= tf) & (tn > 0) & (tn < best)
        best = np.where(ok, tn, best)
        hit = np.where(ok, g, hit)
        nrm = np.where(ok[..., None], nl @ R.T, nrm)
    inten = HEAD_AMBIENT + HEAD_DIFFUSE * np.maximum(0.0, -np.sum(nrm * D, axis=-1))
    for L in LIGHTS:
        d = L[:3] / np.linalg.norm(L[:3])
        inten = inten + L[3] * np.maximum(0.0, -(nrm @ d))
    img = np.zeros((height, width, 3))
    m_hit = hit >= 0
    img[m_hit] = np.minimum(1.0, rgb[hit[m_hit]] * inten[m_hit][:, None])
    return (img * 255.0 + 0.5).astype(np.uint8), hit
