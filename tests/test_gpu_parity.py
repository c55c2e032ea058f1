"""GPU parity tests: the CUDA path, called through the C ABI (libso100_b200.so), against the
fp64 oracle on identical seeded states.  Tolerances (BASELINE.json north_star: 1e-4 relative
in fp32, tighter contact-free; flags exact; rewards within 1e-5) are written next to each check.
"""
import numpy as np
import pytest

import scenarios
from parity_util import contact_errors, gpu_contacts, inject, make_pair, match_contacts, rel_err

pytestmark = pytest.mark.gpu

N = 64
CONTACT_FREE = ("free_space",)


@pytest.mark.parametrize("name", list(scenarios.ALL))
def test_forward_parity(model_blob, name):
    """mj_forward: contact list (geoms, dist, pos, normal), qacc and contact forces."""
    qpos, qvel, ctrl = scenarios.ALL[name](N)
    sim, orc = make_pair(model_blob, N)
    inject(sim, orc, qpos, qvel, ctrl)
    orc.forward()
    fwd = sim.forward()
    qacc_g = fwd["qacc"].cpu().numpy().astype(np.float64)
    worst = dict(dist=0.0, pos=0.0, normal=0.0, force=0.0)
    for i in range(N):
        pairs = match_contacts(gpu_contacts(fwd, i), orc.contacts(i))
        assert pairs is not None, f"env {i}: contact sets differ: gpu={gpu_contacts(fwd, i)} oracle={orc.contacts(i)}"
        e = contact_errors(pairs)
        for k in worst:
            worst[k] = max(worst[k], e[k])
    qacc_o = np.stack([orc.dyn(i)["qacc"] for i in range(N)])
    # geometry: float32 positions of O(0.5 m) -> 1e-6 m; normals 1e-5
    assert worst["dist"] < 2e-6 and worst["pos"] < 2e-5 and worst["normal"] < 2e-5, worst
    # accelerations: arm block ~1e-4 relative; cube rows scale with the contact stiffness
    tol = 2e-4 if name in CONTACT_FREE else 2e-3
    err = rel_err(qacc_g, qacc_o, floor=1.0)
    assert err < tol, (name, err, worst)
    assert worst["force"] < 2e-3, worst
    sites_o = np.stack([orc.dyn(i)["sites"][[3, 4, 2]] for i in range(N)])
    assert np.abs(fwd["sites"].cpu().numpy() - sites_o).max() < 1e-6
    sim.close()


@pytest.mark.parametrize("name", list(scenarios.ALL))
def test_single_substep_parity(model_blob, name):
    """One mj_step from identical qpos/qvel/ctrl: qpos/qvel agreement."""
    qpos, qvel, ctrl = scenarios.ALL[name](N)
    sim, orc = make_pair(model_blob, N)
    inject(sim, orc, qpos, qvel, ctrl)
    orc.substeps(1)
    sim.substeps(1)
    qp_o, qv_o, _, _ = orc.get_state()
    qp_g, qv_g, _, _ = [t.cpu().numpy().astype(np.float64) for t in sim.get_state()]
    tol_v = 1e-5 if name in CONTACT_FREE else 1e-4     # relative to (1 + |v|)
    assert rel_err(qv_g, qv_o, floor=1.0) < tol_v, name
    assert np.abs(qp_g - qp_o).max() < 2e-6, name       # h * dv plus float32 rounding of qpos
    sim.close()


def test_env_step_parity_contact_free(model_blob):
    """Config 2: full env.step (unnormalise + 10 substeps + trailing forward + obs) near the start pose."""
    rng = np.random.default_rng(7)
    n = 64
    sim, orc = make_pair(model_blob, n, task=0, seed=11)
    pose = np.zeros((n, 7)); pose[:, 0] = -0.2; pose[:, 1] = 0.45; pose[:, 2] = 0.3; pose[:, 3] = 1
    import torch
    orc.reset(box_pose=pose)
    sim.reset(box_pose=torch.tensor(pose, dtype=torch.float32))
    a_start = np.array([0, 0.35089, -0.19493, 0, 0, -0.79585])
    for _ in range(3):
        act = (a_start + rng.uniform(-0.1, 0.1, size=(n, 6))).astype(np.float32)
        out_o = orc.step(act, autoreset=False)
        obs, rew, term, trunc, succ = sim.step(torch.tensor(act), autoreset=False)
        assert np.abs(obs.cpu().numpy() - out_o["obs"]).max() < 2e-5
        assert np.array_equal(rew.cpu().numpy(), out_o["reward"])
        assert np.array_equal(term.cpu().numpy().astype(bool), out_o["terminated"])
        assert np.array_equal(trunc.cpu().numpy().astype(bool), out_o["truncated"])
    qp_o, qv_o, ctrl_o, _ = orc.get_state()
    qp_g, qv_g, ctrl_g, _ = [t.cpu().numpy().astype(np.float64) for t in sim.get_state()]
    assert np.abs(ctrl_g - ctrl_o).max() == 0.0          # unnormalize_so100 is bit-exact in float32
    assert np.abs(qp_g[:, :6] - qp_o[:, :6]).max() < 2e-5
    assert rel_err(qv_g[:, :6], qv_o[:, :6]) < 2e-4
    sim.close()


def test_reset_sampling_bit_exact(model_blob):
    """On-device Philox cube placement == the oracle's, bit for bit, and independent of sharding."""
    import torch
    n = 256
    sim, orc = make_pair(model_blob, n, task=1, seed=0x50100, env_offset=1000)
    obs_o, ag_o, dg_o = orc.reset()
    obs_g, ag_g, dg_g = sim.reset()
    qp_g = sim.get_state()[0].cpu().numpy()
    qp_o = orc.get_state()[0]
    assert np.array_equal(qp_g[:, 6:13], qp_o[:, 6:13].astype(np.float32))
    assert np.array_equal(dg_g.cpu().numpy(), dg_o)
    assert np.abs(obs_g.cpu().numpy() - obs_o).max() < 1e-6
    assert (qp_g[:, 6] >= -0.25).all() and (qp_g[:, 6] <= -0.15).all() and (qp_g[:, 7] >= 0.3).all() and (qp_g[:, 7] <= 0.6).all()
    # a shard that starts at global env 1100 reproduces envs 100.. of the first handle
    sim2, _ = make_pair(model_blob, 64, task=1, seed=0x50100, env_offset=1100)
    sim2.reset()
    assert np.array_equal(sim2.get_state()[0].cpu().numpy()[:, 6:9], qp_g[100:164, 6:9])
    sim.close(); sim2.close()


def test_compute_reward_batch_bit_exact(model_blob):
    import torch
    from oracle.so100_oracle import compute_reward
    rng = np.random.default_rng(3)
    ag = rng.uniform(-0.3, 0.3, size=(4096, 3)).astype(np.float32)
    dg = (ag + rng.normal(scale=0.008, size=ag.shape)).astype(np.float32)
    sim, _ = make_pair(model_blob, 4)
    r = sim.compute_reward(torch.tensor(ag), torch.tensor(dg)).cpu().numpy()
    assert np.array_equal(r, compute_reward(ag, dg))
    assert set(np.unique(r)) <= {0.0, -1.0} and (r == 0).any() and (r == -1).any()
    sim.close()
