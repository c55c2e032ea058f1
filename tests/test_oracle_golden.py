"""Pins the oracle (and the host-side Python mirror) on vectors produced by the reference's own
Python code (tests/golden/make_golden.py) and on the reference's test table
(/root/reference/tests/test_constants.py:6-35)."""
import json
import os

import numpy as np
import pytest

from oracle import so100_oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))


@pytest.fixture(scope="module")
def orc(model_blob):
    o = O.Oracle(model_blob, 1)
    o.reset(box_pose=np.array([[-0.2, 0.45, 0.05, 1, 0, 0, 0]]))
    yield o
    o.close()


def test_unnormalize_known_answers():
    """The 11 known answers of the reference's own unit test, through a python restatement of constants.py:44-47."""
    def unnormalize(num, lo, hi, omin=-1, omax=1):
        return float(np.clip((num - omin) / (omax - omin) * (hi - lo) + lo, lo, hi))
    for case in GOLD["unnormalize_known"]:
        assert unnormalize(*case["args"]) == pytest.approx(case["value"], abs=1e-12)
    assert unnormalize(-2, -10, 10) == -10 and unnormalize(2, -10, 10) == 10


def test_unnormalize_so100_bit_exact(orc):
    act = np.array(GOLD["unnormalize_so100"]["action"], dtype=np.float32)
    want = np.array(GOLD["unnormalize_so100"]["ctrl"], dtype=np.float32)
    got = O.unnormalize(orc, act)
    assert np.array_equal(got, want)


def test_box_pose_sampling_matches_reference():
    from gym_so100_c_b200.vec_env import sample_so100_box_pose
    for case in GOLD["box_pose"]:
        assert np.array_equal(sample_so100_box_pose(case["seed"]), np.array(case["pose"]))


def test_compute_reward_matches_reference():
    g = GOLD["compute_reward"]
    ag, dg = np.array(g["achieved"], np.float32), np.array(g["desired"], np.float32)
    got = O.compute_reward(ag, dg)
    assert np.array_equal(got, np.array(g["batch"], np.float32))
    assert np.array_equal(got, np.array(g["single"], np.float32))         # scalar branch == batch branch
    assert np.array_equal(got == 0.0, np.array(g["success"]))


def test_cube_to_bin_reward_truth_table(orc):
    levels = set()
    for case in GOLD["cube_to_bin_reward"]:
        r = O.test_reward(orc, case["contacts"], case["cube_site"])
        assert r == case["reward"], case
        levels.add(r)
    assert levels == {0.0, 1.0, 2.0, 2.5, 3.0, 4.0}


def test_touch_reward_tables(orc):
    """SO100TouchCubeTask / SO100TouchCubeSparseTask.get_reward (single_arm.py:149-215, 246-285) over all distance tiers,
    with and without pad contact: the oracle's float64 restatement equals the reference's float64 result, cast to float32."""
    seen = set()
    for case in GOLD["touch_reward"]:
        task = 2 if case["task"] == "so100_touch_cube" else 3
        r = O.test_touch_reward(orc, task, case["contacts"], case["cube_site"], case["ee_site"])
        assert r == np.float32(case["reward"]), case
        seen.add((case["task"], r == 4.0))
    assert seen == {("so100_touch_cube", True), ("so100_touch_cube", False), ("so100_touch_cube_sparse", True),
                    ("so100_touch_cube_sparse", False)}
    sparse = {np.float32(c["reward"]) for c in GOLD["touch_reward"] if c["task"] == "so100_touch_cube_sparse"}
    assert sparse == {np.float32(-0.2), np.float32(4.0)}
