"""Seeded injected states shared by the oracle tests and the GPU parity tests.

Each generator returns float64 arrays (qpos [N,13], qvel [N,12], ctrl [N,6]) that are first
rounded to float32 so the CUDA path and the fp64 oracle start from bit-identical states.
"""
from __future__ import annotations

import numpy as np

START = np.array([0.0, -0.96, 1.16, 0.0, 0.0, 0.02239])
BIN_XY = np.array([-0.2, 0.7])


def _f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def _rand_quat(rng, n, max_angle):
    ax = rng.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = rng.uniform(-max_angle, max_angle, size=(n, 1))
    return np.concatenate([np.cos(ang / 2), ax * np.sin(ang / 2)], axis=1)


def _base(rng, n, arm_spread=0.1, vel=0.5):
    qpos = np.zeros((n, 13))
    qpos[:, :6] = START + rng.uniform(-arm_spread, arm_spread, size=(n, 6))
    qpos[:, 5] = np.clip(qpos[:, 5], -0.1, 1.7)
    qvel = np.zeros((n, 12))
    qvel[:, :6] = rng.uniform(-vel, vel, size=(n, 6))
    ctrl = START + rng.uniform(-arm_spread, arm_spread, size=(n, 6))
    return qpos, qvel, ctrl


def free_space(n, seed=0):
    """Config 2: arm near the start pose (no hull overlaps, SURVEY 7 hard part 3), cube in free flight."""
    rng = np.random.default_rng(seed)
    qpos, qvel, ctrl = _base(rng, n)
    qpos[:, 6] = rng.uniform(-0.25, -0.15, n)
    qpos[:, 7] = rng.uniform(0.3, 0.6, n)
    qpos[:, 8] = rng.uniform(0.06, 0.12, n)
    qpos[:, 9:13] = _rand_quat(rng, n, np.pi)
    qvel[:, 6:9] = rng.uniform(-0.3, 0.3, size=(n, 3))
    qvel[:, 9:12] = rng.uniform(-2, 2, size=(n, 3))
    return _f32(qpos), _f32(qvel), _f32(ctrl)


def cube_on_table(n, seed=1, tilt=0.3, flat=False):
    """Cube touching / penetrating the table top with a generic (vertex-face) or flat orientation."""
    rng = np.random.default_rng(seed)
    qpos, qvel, ctrl = _base(rng, n)
    qpos[:, 6] = rng.uniform(-0.25, -0.15, n)
    qpos[:, 7] = rng.uniform(0.3, 0.6, n)
    if flat:
        yaw = rng.uniform(-np.pi, np.pi, n)
        quat = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], axis=1)
        qpos[:, 8] = 0.02 - rng.uniform(1e-5, 1e-3, n)
    else:
        quat = _rand_quat(rng, n, tilt)
        # lowest corner of the cube slightly below z = 0
        from gym_so100_c_b200.mjcf import quat_to_mat
        for i in range(n):
            R = quat_to_mat(quat[i])
            corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * 0.02
            zmin = (corners @ R.T)[:, 2].min()
            qpos[i, 8] = -zmin - rng.uniform(1e-5, 2e-3)
    qpos[:, 9:13] = quat
    qvel[:, 6:9] = rng.uniform(-0.2, 0.2, size=(n, 3))
    qvel[:, 8] = rng.uniform(-0.3, 0.05, n)
    qvel[:, 9:12] = rng.uniform(-1, 1, size=(n, 3))
    return _f32(qpos), _f32(qvel), _f32(ctrl)


def cube_in_bin(n, seed=2):
    """Cube resting on the bin floor, some of them pressed into a wall (box-box multi-contact)."""
    rng = np.random.default_rng(seed)
    qpos, qvel, ctrl = _base(rng, n)
    off = rng.uniform(-0.036, 0.036, size=(n, 2))
    off[::3, 0] = 0.0355 + rng.uniform(0, 0.0004, size=len(off[::3]))   # into wall3 (+x)
    qpos[:, 6:8] = BIN_XY + off
    yaw = rng.uniform(-0.05, 0.05, n)
    tilt = _rand_quat(rng, n, 0.02)
    qpos[:, 9:13] = tilt
    qpos[:, 9] = np.cos(yaw / 2) * tilt[:, 0]
    qpos[:, 9:13] /= np.linalg.norm(qpos[:, 9:13], axis=1, keepdims=True)
    qpos[:, 8] = 0.001 + 0.02 - rng.uniform(1e-5, 8e-4, n)    # bin floor top is z = 0.001
    qvel[:, 6:9] = rng.uniform(-0.1, 0.1, size=(n, 3))
    qvel[:, 9:12] = rng.uniform(-0.5, 0.5, size=(n, 3))
    return _f32(qpos), _f32(qvel), _f32(ctrl)


def limits(n, seed=3):
    """Arm joints pushed past their limits (limit rows active), cube in flight."""
    rng = np.random.default_rng(seed)
    qpos, qvel, ctrl = free_space(n, seed)
    lo = np.array([-1.92, -3.32, -0.174, -1.66, -2.79, -0.174])
    hi = np.array([1.92, 0.174, 3.14, 1.66, 2.79, 1.75])
    for i in range(n):
        j = int(rng.integers(3, 6))                  # wrist / jaw joints keep the arm clear of the table
        over = rng.uniform(1e-4, 5e-3)
        qpos[i, j] = hi[j] + over if rng.random() < 0.5 else lo[j] - over
    qpos[:, 5] = np.where(np.arange(n) % 2 == 0, hi[5] + 1e-3, qpos[:, 5])
    qpos[:, 8] = rng.uniform(0.25, 0.3, n)           # cube well above the gripper's reach
    return _f32(qpos), _f32(qvel), _f32(ctrl)


def _site_pose(qarm):
    """ee_site position and Fixed_Jaw frame for one arm configuration (numpy FK of the loader)."""
    from gym_so100_c_b200 import model
    from gym_so100_c_b200.mjcf import quat_to_mat
    m = model.load_model()
    q = m["qpos0"].copy()
    q[:6] = qarm
    xpos, xquat = model.fk(m, q)
    return xpos[8], quat_to_mat(xquat[8]), xquat[8]


def _sample_filtered(n, seed, kind):
    """Rejection-sample states whose ORACLE contact list contains general-hull pairs (GJK/EPA path):
    kind = "arm": random arm configurations that touch the table / base / themselves;
    kind = "grasp": cube dropped into the gripper gap (jaw hulls + pads against the cube)."""
    from gym_so100_c_b200 import model
    from oracle.so100_oracle import Oracle
    m = model.load_model()
    blob = model.pack(m)
    hull_ids = {int(m["geom_mjid"][g]) for g in range(int(m["ngeom"])) if int(m["geom_type"][g]) == 7 and int(m["geom_vnum"][g]) != 8}
    rng = np.random.default_rng(seed)
    lo = np.array([-1.92, -3.32, -0.174, -1.66, -2.79, -0.174]) + 0.02
    hi = np.array([1.92, 0.174, 3.14, 1.66, 2.79, 1.75]) - 0.02
    keep_q, keep_v, keep_c = [], [], []
    batch = 128
    orc = Oracle(blob, batch)
    while len(keep_q) < n:
        qpos = np.zeros((batch, 13)); qvel = np.zeros((batch, 12)); ctrl = np.zeros((batch, 6))
        if kind == "arm":
            qpos[:, :6] = rng.uniform(lo, hi, size=(batch, 6))
            qpos[:, 6:9] = [0.3, 0.3, 0.3]      # cube parked far away
            qpos[:, 9] = 1
            qpos[:, 9:13] = _rand_quat(rng, batch, np.pi)
        else:
            qpos[:, :6] = START + rng.uniform(-0.3, 0.3, size=(batch, 6))
            qpos[:, 5] = rng.uniform(0.1, 1.2, batch)
            for i in range(batch):
                p8, R8, _ = _site_pose(qpos[i, :6])
                local = np.array([rng.uniform(-0.03, 0.03), rng.uniform(-0.1, -0.03), rng.uniform(-0.015, 0.015)])
                qpos[i, 6:9] = p8 + R8 @ local
            qpos[:, 9:13] = _rand_quat(rng, batch, np.pi)
        qvel[:, :6] = rng.uniform(-0.3, 0.3, size=(batch, 6))
        qvel[:, 6:12] = rng.uniform(-0.2, 0.2, size=(batch, 6))
        ctrl[:] = qpos[:, :6] + rng.uniform(-0.05, 0.05, size=(batch, 6))
        qpos, qvel, ctrl = _f32(qpos), _f32(qvel), _f32(ctrl)
        orc.set_state(qpos, qvel, ctrl, np.zeros((batch, 12)))
        orc.forward()
        for i in range(batch):
            cs = orc.contacts(i)
            if not cs or len(cs) > 16:
                continue
            has_hull = any(c["geom1"] in hull_ids or c["geom2"] in hull_ids for c in cs)
            depth = max(-c["dist"] for c in cs)
            if has_hull and 2e-5 < depth < 6e-3 and len(keep_q) < n:
                keep_q.append(qpos[i]); keep_v.append(qvel[i]); keep_c.append(ctrl[i])
    orc.close()
    return np.stack(keep_q), np.stack(keep_v), np.stack(keep_c)


def arm_hull_contacts(n, seed=5):
    return _sample_filtered(n, seed, "arm")


def grasp_hull_contacts(n, seed=6):
    return _sample_filtered(n, seed, "grasp")


ALL = {
    "free_space": free_space,
    "cube_on_table": cube_on_table,
    "cube_flat": lambda n, seed=4: cube_on_table(n, seed, flat=True),
    "cube_in_bin": cube_in_bin,
    "limits": limits,
    "arm_hull_contacts": arm_hull_contacts,
    "grasp_hull_contacts": grasp_hull_contacts,
}
