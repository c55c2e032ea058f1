"""Oracle vs MuJoCo itself -- runs only where `mujoco` (3.3.x) and the reference's assets are available; skipped otherwise.

Neither this container nor the GPU box has MuJoCo, so the oracle's parity with the reference's physics is
"unpinned" in the committed results.  Whoever has `pip install mujoco==3.3.3` and a checkout of gym-so100-c can pin it:

    SO100_REFERENCE_ASSETS=/path/to/gym_so100/assets python -m pytest tests/test_parity_mujoco.py -q

Checks (fp64 oracle against MuJoCo's mj_forward / mj_step on identical qpos / qvel / ctrl):
  * compiled model constants: nq / nv / nu, body masses, dof_M0-derived actuator kv, geom count, the candidate pair count;
  * contact-free forward dynamics: qacc within 1e-6 relative;
  * scenario states (tests/scenarios.py): contact sets (geom pairs), penetration depth within 1e-6 m, normals within 1e-4,
    qacc within 1e-3 relative for box contacts (the contact POINT of a convex-convex pair whose closest features are parallel
    is algorithm dependent, DESIGN.md section 3, so hull scenarios only compare depth / normal / pair sets);
  * ten substeps from the reset state: qpos within 1e-5.
"""
import os

import numpy as np
import pytest

mujoco = pytest.importorskip("mujoco")

import scenarios  # noqa: E402

ASSETS = os.environ.get("SO100_REFERENCE_ASSETS", "/root/reference/gym_so100/assets")
XML = os.path.join(ASSETS, "so100_transfer_cube.xml")
pytestmark = pytest.mark.skipif(not os.path.exists(XML), reason="reference assets not found (set SO100_REFERENCE_ASSETS)")


@pytest.fixture(scope="module")
def mj():
    model = mujoco.MjModel.from_xml_path(XML)
    return model, mujoco.MjData(model)


@pytest.fixture(scope="module")
def orc(model_blob):
    from oracle.so100_oracle import Oracle
    o = Oracle(model_blob, 1)
    o.reset(box_pose=np.array([[-0.2, 0.45, 0.05, 1, 0, 0, 0]]))
    yield o
    o.close()


def _mj_forward(mj, qpos, qvel, ctrl):
    model, data = mj
    mujoco.mj_resetData(model, data)
    data.qpos[:] = qpos
    data.qvel[:] = qvel
    data.ctrl[:] = ctrl
    data.qacc_warmstart[:] = 0
    mujoco.mj_forward(model, data)
    cons = []
    for i in range(data.ncon):
        c = data.contact[i]
        cons.append(dict(geom1=int(c.geom1), geom2=int(c.geom2), dist=float(c.dist), pos=np.array(c.pos), normal=np.array(c.frame[:3])))
    return np.array(data.qacc), cons


def test_model_constants(mj, model_blob):
    from gym_so100_c_b200 import model as M
    model, _ = mj
    m = M.unpack(model_blob)
    assert (model.nq, model.nv, model.nu) == (int(m["nq"]), int(m["nv"]), int(m["nu"])) == (13, 12, 6)
    kv = -model.actuator_biasprm[:, 2]
    assert np.allclose(kv, np.array(m["act_kv"][:6]), rtol=1e-6)
    assert np.allclose(model.dof_armature[:6], np.array(m["dof_armature"][:6]))
    collidable = int(((model.geom_contype != 0) | (model.geom_conaffinity != 0)).sum())
    assert collidable == int(m["ngeom"]) == 25


def test_contact_free_forward(mj, orc):
    qpos, qvel, ctrl = scenarios.ALL["free_space"](16)
    for i in range(16):
        qacc_m, cons = _mj_forward(mj, qpos[i], qvel[i], ctrl[i])
        assert not cons
        orc.set_state(qpos[i:i + 1], qvel[i:i + 1], ctrl[i:i + 1], np.zeros((1, 12)))
        orc.forward()
        qacc_o = orc.dyn(0)["qacc"]
        assert np.abs(qacc_o - qacc_m).max() / (1 + np.abs(qacc_m).max()) < 1e-6


@pytest.mark.parametrize("name", ["cube_on_table", "cube_flat", "cube_in_bin", "limits", "arm_hull_contacts", "grasp_hull_contacts"])
def test_scenario_contacts_and_qacc(mj, orc, name):
    qpos, qvel, ctrl = scenarios.ALL[name](16)
    hull = name in ("arm_hull_contacts", "grasp_hull_contacts", "cube_on_table", "cube_flat")   # the table is a mesh geom
    for i in range(16):
        qacc_m, cons_m = _mj_forward(mj, qpos[i], qvel[i], ctrl[i])
        orc.set_state(qpos[i:i + 1], qvel[i:i + 1], ctrl[i:i + 1], np.zeros((1, 12)))
        orc.forward()
        cons_o = orc.contacts(0)
        assert sorted((c["geom1"], c["geom2"]) for c in cons_m) == sorted((c["geom1"], c["geom2"]) for c in cons_o)
        for cm in cons_m:
            same = [co for co in cons_o if (co["geom1"], co["geom2"]) == (cm["geom1"], cm["geom2"])]
            best = min(same, key=lambda co: np.linalg.norm(co["pos"] - cm["pos"]))
            assert abs(min(co["dist"] for co in same) - min(c["dist"] for c in cons_m if (c["geom1"], c["geom2"]) == (cm["geom1"], cm["geom2"]))) < 1e-6
            if not hull:
                assert np.abs(best["normal"] - cm["normal"]).max() < 1e-4
        if not hull:
            qacc_o = orc.dyn(0)["qacc"]
            assert np.abs(qacc_o - qacc_m).max() / (1 + np.abs(qacc_m).max()) < 1e-3


def test_ten_substeps_from_reset(mj, orc):
    model, data = mj
    pose = np.array([[-0.2, 0.45, 0.05, 1, 0, 0, 0]])
    orc.reset(box_pose=pose)
    qpos, qvel, ctrl, _ = orc.get_state()
    mujoco.mj_resetData(model, data)
    data.qpos[:] = qpos[0]
    data.ctrl[:] = ctrl[0]
    mujoco.mj_forward(model, data)
    for _ in range(10):
        mujoco.mj_step(model, data)
    orc.substeps(10)
    qp_o = orc.get_state()[0][0]
    assert np.abs(qp_o - data.qpos).max() < 1e-5
