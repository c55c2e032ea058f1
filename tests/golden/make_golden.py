#!/usr/bin/env python
"""Generate tests/golden/reference_golden.json by EXECUTING the reference's own Python code
(/root/reference, numpy-only parts) in this container.  The GPU box has no /root/reference, so
the vectors are committed; rerun this script to refresh them:

    python tests/golden/make_golden.py [--reference /root/reference]

What is pinned (reference file:line):
  * constants.unnormalize known answers                      tests/test_constants.py:6-35
  * constants.unnormalize_so100 on float32 actions           gym_so100/constants.py:78-86
  * utils.sample_so100_box_pose(seed)                        gym_so100/utils.py:18-29
  * SO100GoalEnv.compute_reward / _is_success                gym_so100/env.py:341-358
  * SO100CubeToBinTask.get_reward on synthetic physics       gym_so100/tasks/single_arm.py:322-380
  * SO100TouchCubeTask / SO100TouchCubeSparseTask.get_reward  gym_so100/tasks/single_arm.py:149-215, 246-285
MuJoCo / dm_control / gymnasium are not installed, so they are replaced by inert stub modules that
only let the reference modules import; the functions above never touch them.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import types

import numpy as np


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Env:
        def __init__(self, *a, **k):
            pass

        def reset(self, seed=None, options=None):
            return None

    class _Space:
        def __init__(self, *a, **k):
            pass

    spaces = mod("gymnasium.spaces", Box=_Space, Dict=_Space)
    reg = mod("gymnasium.envs.registration", register=lambda **k: None)
    envs = mod("gymnasium.envs", registration=reg)
    mod("gymnasium", Env=_Env, spaces=spaces, envs=envs)
    mod("gym")

    class _Task:
        def __init__(self, random=None):
            self._random = random

        def before_step(self, action, physics):
            physics.set_control(action)

        def initialize_episode(self, physics):
            pass

    base = mod("dm_control.suite.base", Task=_Task)
    suite = mod("dm_control.suite", base=base)
    mj = mod("dm_control.mujoco")
    control = mod("dm_control.rl.control")
    rl = mod("dm_control.rl", control=control)
    mod("dm_control", mujoco=mj, rl=rl, suite=suite)


class FakePhysics:
    """Just enough of dm_control's Physics for SO100CubeToBinTask.get_reward."""
    GEOMS = {0: "table", 20: "fixed_jaw_pad_1", 21: "fixed_jaw_pad_2", 22: "fixed_jaw_pad_3", 23: "fixed_jaw_pad_4",
             28: "moving_jaw_pad_1", 29: "moving_jaw_pad_2", 30: "moving_jaw_pad_3", 31: "moving_jaw_pad_4",
             32: "red_box", 33: "bin_wall", 34: "bin_wall2", 35: "bin_wall3", 36: "bin_wall4", 37: "bin_floor",
             18: "", 19: "", 15: ""}
    SITES = {"midair": 0, "left_cam_focus": 1, "ee_site": 2, "cube_site": 3, "bin_center": 4}

    def __init__(self, contacts, cube_site, ee_site=(-0.05, 0.5, 0.15), bin_center=(-0.2, 0.7, 0.021)):
        outer = self

        class _Model:
            def site(self, name):
                return types.SimpleNamespace(id=outer.SITES[name])

            def id2name(self, i, kind):
                return outer.GEOMS[int(i)]

        xpos = np.zeros((5, 3))
        xpos[2], xpos[3], xpos[4] = ee_site, cube_site, bin_center
        self.model = _Model()
        self.data = types.SimpleNamespace(
            site_xpos=xpos, ncon=len(contacts),
            contact=[types.SimpleNamespace(geom1=a, geom2=b) for a, b in contacts])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    install_stubs()
    with contextlib.redirect_stdout(io.StringIO()):
        from gym_so100 import constants, utils
        from gym_so100.env import SO100GoalEnv
        from gym_so100.tasks.single_arm import SO100CubeToBinTask, SO100TouchCubeSparseTask, SO100TouchCubeTask

    out = {"numpy": np.__version__}

    # tests/test_constants.py:6-35
    ka = [(-1, -10, 10), (1, -10, 10), (0, -10, 10), (0.5, -10, 10), (-0.5, -10, 10), (-2, -10, 10), (2, -10, 10),
          (0, 0, 20), (-1, 0, 20), (1, 0, 20), (0.25, -5.0, 5.0)]
    out["unnormalize_known"] = [dict(args=list(a), value=float(constants.unnormalize(*a))) for a in ka]

    rng = np.random.RandomState(20240607)
    acts = np.concatenate([
        np.zeros((1, 6)), np.ones((1, 6)), -np.ones((1, 6)), np.array([[0.5, -0.5, 0.25, -0.25, 2, -2]]),
        rng.uniform(-1, 1, size=(60, 6)), rng.uniform(-1.5, 1.5, size=(12, 6))]).astype(np.float32)
    un = np.stack([constants.unnormalize_so100(a.copy()) for a in acts])
    assert un.dtype == np.float32
    out["unnormalize_so100"] = dict(action=acts.tolist(), ctrl=un.tolist())

    seeds = [0, 1, 2, 3, 7, 41, 123, 1234, 99999, 2**31 - 1]
    out["box_pose"] = [dict(seed=s, pose=utils.sample_so100_box_pose(s).tolist()) for s in seeds]

    # env.py:341-358 with a stand-in `self`
    dummy = types.SimpleNamespace(distance_threshold=0.01)
    ag = rng.uniform(-0.3, 0.3, size=(64, 3)).astype(np.float32)
    dg = (ag + rng.normal(scale=0.007, size=ag.shape)).astype(np.float32)
    dg[0] = ag[0]
    dg[1] = ag[1] + np.array([0.01, 0, 0], dtype=np.float32)
    batch = SO100GoalEnv.compute_reward(dummy, ag, dg, {})
    single = [float(SO100GoalEnv.compute_reward(dummy, ag[i], dg[i], {})) for i in range(len(ag))]
    succ = [bool(SO100GoalEnv._is_success(dummy, ag[i], dg[i])) for i in range(len(ag))]
    out["compute_reward"] = dict(achieved=ag.tolist(), desired=dg.tolist(), batch=batch.tolist(), single=single, success=succ)

    # single_arm.py:322-380 truth table
    task = SO100CubeToBinTask()
    cases = []
    cube_sites = {
        "table": (-0.19, 0.46, 0.03), "over_bin_high": (-0.2, 0.7, 0.2), "inside": (-0.2, 0.7, 0.0351),
        "inside_edge_x": (-0.2502, 0.7, 0.0351), "too_low": (-0.2, 0.7, 0.0309), "too_high": (-0.2, 0.7, 0.0411),
        "just_inside_low": (-0.2, 0.7, 0.03101), "xy_edge_in": (-0.2599, 0.6401, 0.035), "xy_edge_out": (-0.2601, 0.7, 0.035),
    }
    contact_sets = {
        "none": [], "table": [(32, 0)], "table_rev": [(0, 32)], "pad": [(20, 32)], "pad_rev": [(32, 30)],
        "pad+table": [(23, 32), (32, 0)], "pad+table_rev": [(23, 32), (0, 32)], "bin": [(32, 37), (32, 35)],
        "bin+pad": [(32, 37), (29, 32)], "jaw_hull": [(32, 18)], "two_pads": [(21, 32), (28, 32), (32, 36)],
    }
    for sname, site in cube_sites.items():
        for cname, cons in contact_sets.items():
            with contextlib.redirect_stdout(io.StringIO()):
                r = task.get_reward(FakePhysics(cons, site))
            cases.append(dict(site=sname, cube_site=list(site), contacts=[list(c) for c in cons], reward=float(r)))
    out["cube_to_bin_reward"] = cases

    # single_arm.py:149-215 (shaped) and 246-285 (sparse): distance tiers x pad contact
    tcases = []
    cube = np.array([-0.2, 0.45, 0.03])
    direction = np.array([0.6, -0.48, 0.64])
    dists = [0.9, 0.7, 0.6999, 0.55, 0.5, 0.42, 0.3, 0.2999, 0.17, 0.1, 0.0999, 0.07, 0.05, 0.0499, 0.031, 0.004, 0.0]
    tsets = {"none": [], "table": [(32, 0)], "pad": [(20, 32)], "pad_rev": [(32, 30)], "jaw_hull": [(32, 18)],
             "pad+table": [(23, 32), (32, 0)]}
    for tname, cls in (("so100_touch_cube", SO100TouchCubeTask), ("so100_touch_cube_sparse", SO100TouchCubeSparseTask)):
        ttask = cls()
        for d in dists:
            ee = cube + direction * d
            for cname, cons in tsets.items():
                with contextlib.redirect_stdout(io.StringIO()):
                    r = ttask.get_reward(FakePhysics(cons, tuple(cube), ee_site=tuple(ee)))
                tcases.append(dict(task=tname, dist=d, cube_site=cube.tolist(), ee_site=ee.tolist(),
                                   contacts=[list(c) for c in cons], reward=float(r)))
    out["touch_reward"] = tcases

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {path}: {len(out['unnormalize_so100']['action'])} actions, {len(out['box_pose'])} poses, "
          f"{len(ag)} goals, {len(cases)} + {len(tcases)} reward cases (numpy {np.__version__})")


if __name__ == "__main__":
    main()
