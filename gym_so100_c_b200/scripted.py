"""Scripted pick-and-place actions for the bin-a-cube scene (SURVEY.md 8d "Synthetic inputs", BASELINE north_star
"synthetic random and scripted actions").

The reference has no scripted policy (its demonstrations come from teleoperation, scripts/record_teleop.py); this is the
synthetic counterpart for benchmarks and behavioural tests: per env a waypoint sequence in normalised joint space

    start pose -> open gripper -> above the cube -> down around the cube -> close -> lift -> above the bin -> open

with linear interpolation between waypoints, 300-step episodes (env.py:200).  Joint targets come from an inverse-kinematics
table over the cube's reset range (utils.py:19-28: x in [-0.25, -0.15], y in [0.3, 0.6]) built once per process with the numpy
kinematics of model.py and interpolated bilinearly per env.  Host-side numpy only; the actions are ordinary inputs of
`so100_step`.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Tuple

import numpy as np

from . import model as _model

# grasp point in the Fixed_Jaw frame: between the fixed pads (local x = +0.0116 surface) and the moving pads, two thirds down
# the pad row; a 4 cm cube held against the fixed pads is centred here
GRASP_LOCAL = np.array([-0.022, -0.092, 0.0])
JAW_OPEN = 1.2        # rad: pad gap about 8 cm
JAW_CLOSED = 0.15     # rad: commanded beyond contact so that the position actuator keeps squeezing
HOVER_Z = 0.10        # grasp point height above the table while travelling
GRASP_Z = 0.028       # grasp point height when closing: jaw tips (1.4 cm further down) stay clear of the table
BIN_XY = (-0.2, 0.7)  # bin_center site, so100_transfer_cube.xml:16,23
LIFT_Z = 0.14         # grasp point height while carrying: the bin walls are 6.1 cm high, the held cube hangs 2-3 cm below the point
BIN_Z = 0.12          # grasp point height over the bin when the gripper opens
FJ_BODY = 8           # Fixed_Jaw body id


def normalize_so100(m: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Inverse of constants.unnormalize_so100 (constants.py:49-57): joint targets -> actions in [-1, 1]."""
    lo = np.asarray(m["act_lo"][:6], dtype=np.float64)
    hi = np.asarray(m["act_hi"][:6], dtype=np.float64)
    return 2.0 * (np.asarray(q, dtype=np.float64) - lo) / (hi - lo) - 1.0


def _quat_mul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], axis=-1)


def _quat_rot(q, v):
    w, u = q[..., :1], q[..., 1:]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def _grasp_frame(m: np.ndarray, arm5: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Grasp point and the Fixed_Jaw +y axis in the world for a batch of arm poses [..., 5]: the kinematic chain world ->
    Fixed_Jaw of the packed model (body_parent / body_pos / body_quat / body_jaxis), vectorised over the batch."""
    arm5 = np.asarray(arm5, dtype=np.float64)
    chain = []
    b = FJ_BODY
    while b > 0:
        chain.append(b)
        b = int(m["body_parent"][b])
    pos = np.zeros(arm5.shape[:-1] + (3,))
    quat = np.zeros(arm5.shape[:-1] + (4,))
    quat[..., 0] = 1.0
    for b in reversed(chain):
        pos = pos + _quat_rot(quat, np.asarray(m["body_pos"][b], dtype=np.float64))
        quat = _quat_mul(quat, np.asarray(m["body_quat"][b], dtype=np.float64))
        if int(m["body_jtype"][b]) == _model.JNT_HINGE:
            a = int(m["body_qposadr"][b])
            half = 0.5 * (arm5[..., a] - float(m["qpos0"][a]))
            ax = np.asarray(m["body_jaxis"][b], dtype=np.float64)
            jq = np.concatenate([np.cos(half)[..., None], np.sin(half)[..., None] * ax], axis=-1)
            quat = _quat_mul(quat, jq)
    return pos + _quat_rot(quat, GRASP_LOCAL), _quat_rot(quat, np.array([0.0, 1.0, 0.0]))


AXIS_WEIGHT = 0.05    # metres per unit of approach-axis error in the IK residual


def solve_ik(m: np.ndarray, target: np.ndarray, q0: np.ndarray, iters: int = 80) -> Tuple[np.ndarray, np.ndarray]:
    """Arm joints [..., 5] that put the grasp point at `target` [..., 3] with the gripper pointing straight down (Fixed_Jaw
    local +y = world +z): batched Levenberg-Marquardt on finite-difference Jacobians, joint ranges enforced by clipping.
    Returns (q, residual norm)."""
    lo = np.asarray(m["dof_range"][:5, 0], dtype=np.float64) + 1e-3
    hi = np.asarray(m["dof_range"][:5, 1], dtype=np.float64) - 1e-3
    up = np.array([0.0, 0.0, 1.0])
    target = np.asarray(target, dtype=np.float64)

    def resid(q):
        p, y = _grasp_frame(m, q)
        return np.concatenate([p - target, AXIS_WEIGHT * (y - up)], axis=-1)

    q = np.clip(np.broadcast_to(np.asarray(q0, dtype=np.float64), target.shape[:-1] + (5,)).copy(), lo, hi)
    r = resid(q)
    lam = np.full(q.shape[:-1] + (1, 1), 1e-3)
    eye = np.eye(5)
    for _ in range(iters):
        J = np.zeros(r.shape + (5,))
        for k in range(5):
            dq = np.zeros(5)
            dq[k] = 1e-6
            J[..., k] = (resid(q + dq) - r) / 1e-6
        Jt = np.swapaxes(J, -1, -2)
        step = np.linalg.solve(Jt @ J + lam * eye, -(Jt @ r[..., None]))[..., 0]
        qn = np.clip(q + np.clip(step, -0.3, 0.3), lo, hi)
        rn = resid(qn)
        better = (rn * rn).sum(-1) < (r * r).sum(-1)
        q = np.where(better[..., None], qn, q)
        r = np.where(better[..., None], rn, r)
        lam = np.where(better[..., None, None], np.maximum(lam * 0.5, 1e-7), lam * 4.0)
    return q, np.sqrt((r * r).sum(-1))


@lru_cache(maxsize=2)
def _ik_table(blob: bytes):
    """IK solutions (hover and grasp heights) on a 1 cm grid over the cube's reset range, plus the hover pose over the bin."""
    m = _model.unpack(blob)
    xs = np.linspace(float(m["box_lo"][0]), float(m["box_hi"][0]), 11)
    ys = np.linspace(float(m["box_lo"][1]), float(m["box_hi"][1]), 31)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    seed = np.array([0.0, -1.3, 1.3, 1.5, 0.0])
    hover, r1 = solve_ik(m, np.stack([gx, gy, np.full_like(gx, HOVER_Z)], axis=-1), seed)
    grasp, r2 = solve_ik(m, np.stack([gx, gy, np.full_like(gx, GRASP_Z)], axis=-1), hover)
    lift, r3 = solve_ik(m, np.stack([gx, gy, np.full_like(gx, LIFT_Z)], axis=-1), hover)
    bin_hi, r4 = solve_ik(m, np.array([BIN_XY[0], BIN_XY[1], LIFT_Z]), lift[5, -1])
    bin_lo, r5 = solve_ik(m, np.array([BIN_XY[0], BIN_XY[1], BIN_Z]), bin_hi)
    return m, xs, ys, {"hover": hover, "grasp": grasp, "lift": lift, "bin_hi": bin_hi, "bin_lo": bin_lo}, \
        float(max(r1.max(), r2.max(), r3.max(), r4, r5))


# keyframes (episode step, arm pose, jaw): the commanded target is interpolated linearly between consecutive keyframes
_KEYS = ((0, "start", "rest"), (20, "start", "open"), (80, "hover", "open"), (125, "grasp", "open"), (150, "grasp", "closed"),
         (185, "lift", "closed"), (245, "bin_hi", "closed"), (265, "bin_lo", "closed"), (280, "bin_lo", "open"))
CUBE_SITE_OFFSET = 0.01   # obs[:, 0:2] is the cube_site position, (0.01, 0.01, 0.01) off the cube centre (so100_transfer_cube.xml:13)


class ScriptedPolicy:
    """The script as a batched open-loop policy on torch tensors (any device): `reset` fixes every env's keyframes from its
    cube position, `step` returns the actions [N, 6] of the current episode step and advances it.  Episodes restart every
    `period` steps or when `observe` sees an env finish."""

    def __init__(self, model_blob: bytes, num_envs: int, device="cpu", period: int = 300):
        import torch
        self.torch = torch
        m, xs, ys, tab, _ = _ik_table(bytes(model_blob))
        dev = torch.device(device)
        f64 = dict(dtype=torch.float64, device=dev)
        self.n, self.device, self.period = num_envs, dev, period
        self.x0, self.dx, self.nx = float(xs[0]), float(xs[1] - xs[0]), xs.size
        self.y0, self.dy, self.ny = float(ys[0]), float(ys[1] - ys[0]), ys.size
        self.grid = {k: torch.tensor(tab[k], **f64) for k in ("hover", "grasp", "lift")}
        self.fixed = {"start": torch.tensor(np.asarray(m["start_pose"][:5], dtype=np.float64), **f64),
                      "bin_hi": torch.tensor(tab["bin_hi"], **f64), "bin_lo": torch.tensor(tab["bin_lo"], **f64)}
        self.jaw = {"rest": float(m["start_pose"][5]), "open": JAW_OPEN, "closed": JAW_CLOSED}
        self.lo = torch.tensor(np.asarray(m["act_lo"][:6], dtype=np.float64), **f64)
        self.hi = torch.tensor(np.asarray(m["act_hi"][:6], dtype=np.float64), **f64)
        self.times = torch.tensor([float(k[0]) for k in _KEYS], **f64)
        self.keys = torch.zeros((num_envs, len(_KEYS), 6), **f64)
        self.t = torch.zeros(num_envs, dtype=torch.int64, device=dev)

    def _bilinear(self, table, xy):
        torch = self.torch
        fx = torch.clamp((xy[:, 0] - self.x0) / self.dx, 0, self.nx - 1 - 1e-9)
        fy = torch.clamp((xy[:, 1] - self.y0) / self.dy, 0, self.ny - 1 - 1e-9)
        i, j = fx.long(), fy.long()
        a, b = (fx - i)[:, None], (fy - j)[:, None]
        return ((1 - a) * (1 - b) * table[i, j] + a * (1 - b) * table[i + 1, j] + (1 - a) * b * table[i, j + 1]
                + a * b * table[i + 1, j + 1])

    def reset(self, box_xy, mask=None):
        """New episode for the envs in `mask` (all when None) with cube centres `box_xy` [N, 2]."""
        torch = self.torch
        xy = torch.as_tensor(box_xy, dtype=torch.float64, device=self.device).reshape(self.n, 2)
        poses = {k: self._bilinear(g, xy) for k, g in self.grid.items()}
        rows = []
        for _, p, j in _KEYS:
            arm = poses[p] if p in poses else self.fixed[p].expand(self.n, 5)
            rows.append(torch.cat([arm, torch.full((self.n, 1), self.jaw[j], dtype=torch.float64, device=self.device)], dim=1))
        keys = torch.stack(rows, dim=1)
        if mask is None:
            self.keys, self.t = keys, torch.zeros_like(self.t)
        else:
            mk = torch.as_tensor(mask, device=self.device).bool()
            self.keys = torch.where(mk[:, None, None], keys, self.keys)
            self.t = torch.where(mk, torch.zeros_like(self.t), self.t)

    def observe(self, obs, done):
        """After an env step: envs that finished (auto-reset, `obs` already shows the new episode) or whose script has run
        out start over from the cube position in `obs`."""
        torch = self.torch
        o = torch.as_tensor(obs, device=self.device)
        restart = torch.as_tensor(done, device=self.device).bool() | (self.t >= self.period)
        self.reset(o[:, 0:2].double() - CUBE_SITE_OFFSET, restart)

    def step(self):
        torch = self.torch
        t = (self.t + 1).double()                      # the action of a step commands where the arm should be at the end of it
        k = torch.clamp(torch.searchsorted(self.times, t, right=True) - 1, 0, len(_KEYS) - 2)
        t0, t1 = self.times[k], self.times[k + 1]
        w = torch.clamp((t - t0) / (t1 - t0), 0.0, 1.0)[:, None]
        idx = k[:, None, None].expand(self.n, 1, 6)
        q = (1 - w) * self.keys.gather(1, idx)[:, 0] + w * self.keys.gather(1, idx + 1)[:, 0]
        self.t = self.t + 1
        return torch.clamp(2.0 * (q - self.lo) / (self.hi - self.lo) - 1.0, -1.0, 1.0).float()


def scripted_actions(model_blob: bytes, box_xy: np.ndarray, n_steps: int = 300) -> np.ndarray:
    """Actions [n_steps, N, 6] (float32, normalised) of one scripted episode for cubes resting at `box_xy` [N, 2]."""
    xy = np.asarray(box_xy, dtype=np.float64).reshape(-1, 2)
    pol = ScriptedPolicy(model_blob, xy.shape[0], device="cpu", period=10 ** 9)
    pol.reset(xy)
    return np.stack([pol.step().numpy() for _ in range(n_steps)])


def ik_residual(model_blob: bytes) -> float:
    """Largest residual (m / scaled axis error) over the IK table: a check that every grid point is reachable top-down."""
    return _ik_table(bytes(model_blob))[4]
