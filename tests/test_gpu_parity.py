"""GPU parity tests: the CUDA path, called through the C ABI (libso100_b200.so), against the
fp64 oracle on identical seeded states.  Tolerances (BASELINE.json north_star: 1e-4 relative
in fp32, tighter contact-free; flags exact; rewards within 1e-5) are written next to each check.

Scenario classes (tests/scenarios.py):
  * contact-free / box contacts (cube-table, cube-bin, pads): every env must agree;
  * general-hull contacts (GJK/EPA): penetration depth and normal must agree for every env; the
    contact POINT is not unique when two features are parallel (hard part 2),
    and which vertices count as "the deepest feature" flips with the 1e-4 rad resolution of the
    float32 EPA normal, so point / force / acceleration agreement is required for >= 95 % of the envs (measured on
    256 envs per scenario: 97.7-100 %; the rest are parallel-feature contacts where the 1e-6 m deepest-feature band
    sees a different vertex set through a normal that differs by 2e-5 rad).
  * the bench's own state distribution (config 3 after 60 steps of U(-1,1) actions): test_config3_distribution_parity.
"""
import numpy as np
import pytest

import scenarios
from parity_util import contact_errors, gpu_contacts, inject, make_pair, match_contacts, rel_err

pytestmark = pytest.mark.gpu

N = 64
CONTACT_FREE = ("free_space",)
HULL = ("arm_hull_contacts", "grasp_hull_contacts")


def _per_env_forward(model_blob, name):
    qpos, qvel, ctrl = scenarios.ALL[name](N)
    sim, orc = make_pair(model_blob, N)
    inject(sim, orc, qpos, qvel, ctrl)
    orc.forward()
    fwd = sim.forward()
    qacc_g = fwd["qacc"].cpu().numpy().astype(np.float64)
    qacc_o = np.stack([orc.dyn(i)["qacc"] for i in range(N)])
    errs = []
    for i in range(N):
        pairs = match_contacts(gpu_contacts(fwd, i), orc.contacts(i))
        assert pairs is not None, f"env {i}: contact sets differ: gpu={gpu_contacts(fwd, i)} oracle={orc.contacts(i)}"
        e = contact_errors(pairs)
        e["qacc"] = rel_err(qacc_g[i], qacc_o[i], floor=1.0)
        errs.append(e)
    sites_o = np.stack([orc.dyn(i)["sites"][[3, 4, 2]] for i in range(N)])
    assert np.abs(fwd["sites"].cpu().numpy() - sites_o).max() < 1e-6
    sim.close()
    orc.close()
    return errs


@pytest.mark.parametrize("name", list(scenarios.ALL))
def test_forward_parity(model_blob, name):
    """mj_forward: contact list (geoms, dist, pos, normal), qacc and contact forces."""
    errs = _per_env_forward(model_blob, name)
    worst = {k: max(e[k] for e in errs) for k in errs[0]}
    # Tolerances = about 3x the worst case measured on B200 (tests/dev/gpu_tolerance_report.py, 64 envs per scenario):
    #   penetration depth 1.8e-7 m (float32 positions of O(0.5 m)); normals exact for box faces, 8e-5 from the float32 EPA;
    #   contact point 4e-7 m; contact force relative to max(1 N, |f|): 3.3e-4 box contacts, 2.2e-3 grasps (stiff pad contacts, solimp
    #   0.9999); qacc relative to (1 + |qacc|): 4e-6 contact-free, 2.9e-4 box contacts / limits, 1.3e-3 grasps
    assert worst["dist"] < 5e-7, worst
    assert worst["normal"] < (2e-4 if name in HULL else 2e-5), worst
    tol_q = 2e-5 if name in CONTACT_FREE else (2e-3 if name in HULL else 6e-4)
    tol_f = 3e-3 if name in HULL else 1e-3
    ok = [e["pos"] < 2e-6 and e["force"] < tol_f and e["qacc"] < tol_q for e in errs]
    if name in HULL:
        assert np.mean(ok) >= 0.95, (name, float(np.mean(ok)), worst)
    else:
        assert all(ok), (name, worst)


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
    # relative to (1 + |v|); measured worst cases 3.4e-8 contact-free, 1.9e-5 box contacts / limits, 6.7e-5 grasps
    tol_v = 2e-7 if name in CONTACT_FREE else (1e-4 if name in HULL else 5e-5)
    ev = np.array([rel_err(qv_g[i], qv_o[i], floor=1.0) for i in range(N)])
    ep = np.abs(qp_g - qp_o).max(axis=1)                # h * dv plus float32 rounding of qpos
    ok = (ev < tol_v) & (ep < 2e-6)
    if name in HULL:
        assert ok.mean() >= 0.95, (name, float(ok.mean()), float(ev.max()))
    else:
        assert ok.all(), (name, float(ev.max()), float(ep.max()))
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


def test_env_step_parity_cube_landing(model_blob):
    """Config 3 slice: reset, then env steps while the cube falls, lands and settles on the table (box contacts
    through the full phase pipeline): observations within 1e-4 after 8 env steps = 80 substeps, rewards/flags exact."""
    import torch
    rng = np.random.default_rng(9)
    n = 64
    sim, orc = make_pair(model_blob, n, task=0, seed=5)
    orc.reset()
    sim.reset()
    a_start = np.array([0, 0.35089, -0.19493, 0, 0, -0.79585])
    for k in range(8):
        act = (a_start + rng.uniform(-0.05, 0.05, size=(n, 6))).astype(np.float32)
        out_o = orc.step(act, autoreset=False)
        obs, rew, term, trunc, succ = sim.step(torch.tensor(act), autoreset=False)
        assert np.abs(obs.cpu().numpy() - out_o["obs"]).max() < 1e-4, k
        assert np.array_equal(rew.cpu().numpy(), out_o["reward"]), k
        assert np.array_equal(term.cpu().numpy().astype(bool), out_o["terminated"])
    d = sim.diagnostics()
    assert d["contacts_seen"] > 0 and d["nonfinite_resets"] == 0 and d["contact_overflow"] == 0
    sim.close()


def test_reward_levels_on_injected_states(model_blob):
    """Staged reward (single_arm.py:322-380) through real collision: cube on the table (0), cube above the bin (2.5),
    cube inside the bin volume without gripper contact (4 = success -> terminated); all flags exact vs the oracle.
    (A cube RESTING on the bin floor only scores 2.5 in the reference too: its lower face sits a penetration depth
    below the strict `lower > bin_min.z` bound -- SURVEY.md section 7, hard part 5 -- so the 4-case drops it in.)"""
    import torch
    n = 6
    sim, orc = make_pair(model_blob, n, task=0, seed=1)
    orc.reset(); sim.reset()
    qpos, qvel, ctrl, _ = orc.get_state()
    qvel[:] = 0
    cube = np.array([[-0.2, 0.45, 0.01995], [-0.2, 0.45, 0.01995], [-0.2, 0.7, 0.15], [-0.2, 0.7, 0.15],
                     [-0.2, 0.7, 0.026], [-0.21, 0.69, 0.026]])
    qpos[:, 6:9] = cube
    qpos[:, 9:13] = [1, 0, 0, 0]
    inject(sim, orc, scenarios._f32(qpos), qvel, scenarios._f32(ctrl))
    act = np.tile(np.array([0, 0.35089, -0.19493, 0, 0, -0.79585], dtype=np.float32), (n, 1))
    out_o = orc.step(act, autoreset=False)
    obs, rew, term, trunc, succ = sim.step(torch.tensor(act), autoreset=False)
    r = rew.cpu().numpy()
    assert np.array_equal(r, out_o["reward"]), (r, out_o["reward"])
    assert np.array_equal(term.cpu().numpy().astype(bool), out_o["terminated"])
    assert np.array_equal(succ.cpu().numpy().astype(bool), out_o["success"])
    assert list(r[:2]) == [0.0, 0.0] and list(r[2:4]) == [2.5, 2.5] and list(r[4:]) == [4.0, 4.0]
    sim.close()


@pytest.mark.parametrize("task", [2, 3])
def test_touch_tasks_reward_parity(model_blob, task):
    """SO100TouchCube (2) / SO100TouchCubeSparse (3): shaped distance tiers + pad-contact bonus, success = pad contact and
    |ee - cube| < 0.05 (single_arm.py:149-215, 246-285), TimeLimit 300 (__init__.py:7,17).  States: the grasp scenario
    (pads on the cube), arm-hull contacts and free space, so that every reward branch is visited; the reward must agree
    with the float64 oracle within 1e-5 and the flags exactly."""
    import torch
    n = 64
    rewards = []
    for name in ("grasp_hull_contacts", "free_space", "cube_on_table"):
        qpos, qvel, ctrl = scenarios.ALL[name](n)
        sim, orc = make_pair(model_blob, n, task=task, seed=4)
        orc.reset(); sim.reset()
        inject(sim, orc, qpos, np.zeros_like(qvel), ctrl)
        from oracle.so100_oracle import unnormalize  # noqa: F401
        # hold the injected arm pose: action = normalised ctrl
        m = orc_model_ranges(model_blob)
        act = np.clip(2 * (ctrl - m[0]) / (m[1] - m[0]) - 1, -1, 1).astype(np.float32)
        steps = torch.full((n,), 298, dtype=torch.int32)
        sim.set_aux(step_count=steps); orc.set_counters(step_count=steps.numpy())
        for k in range(2):
            out_o = orc.step(act, autoreset=False)
            obs, rew, term, trunc, succ = sim.step(torch.tensor(act), autoreset=False)
            r_g, r_o = rew.cpu().numpy(), out_o["reward"]
            flip = (r_g == 4.0) != (r_o == 4.0)          # knife-edge success threshold (0.05 m) on float32 vs float64 sites
            assert flip.mean() <= 0.05
            assert np.abs(r_g - r_o)[~flip].max() < 1e-5, (name, k)
            assert np.array_equal(term.cpu().numpy().astype(bool)[~flip], out_o["terminated"][~flip])
            assert np.array_equal(trunc.cpu().numpy().astype(bool), out_o["truncated"])
            assert bool(trunc.cpu().numpy().all()) == (k == 1)     # 300-step TimeLimit
            rewards.append(r_g.copy())
        sim.close(); orc.close()
    rewards = np.concatenate(rewards)
    if task == 3:
        assert set(np.unique(rewards)) <= {np.float32(-0.2), np.float32(4.0)}
    else:
        assert rewards.min() >= -0.2 - 1e-6 and len(np.unique(np.round(rewards, 3))) > 10    # many shaped values


def orc_model_ranges(model_blob):
    from gym_so100_c_b200 import model
    m = model.unpack(model_blob)
    return np.array(m["act_lo"][:6], dtype=np.float64), np.array(m["act_hi"][:6], dtype=np.float64)


def test_autoreset_and_truncation(model_blob):
    """TimeLimit semantics: GoalEnv truncates at 300 steps (env.py:395-403); with autoreset the returned obs is the first
    observation of the next episode, final_obs the terminal one, and the episode counter advances."""
    import torch
    n = 32
    sim, orc = make_pair(model_blob, n, task=1, seed=3)
    orc.reset(); sim.reset()
    steps = torch.full((n,), 298, dtype=torch.int32)
    sim.set_aux(step_count=steps)
    orc.set_counters(step_count=steps.numpy())
    act = np.tile(np.array([0, 0.35089, -0.19493, 0, 0, -0.79585], dtype=np.float32), (n, 1))
    for k in range(2):
        out_o = orc.step(act, autoreset=True)
        obs, rew, term, trunc, succ = sim.step(torch.tensor(act), autoreset=True)
        assert np.array_equal(trunc.cpu().numpy().astype(bool), out_o["truncated"])
        assert bool(trunc.cpu().numpy().all()) == (k == 1)
        assert np.array_equal(rew.cpu().numpy(), out_o["reward"])
        assert np.abs(obs.cpu().numpy() - out_o["obs"]).max() < 2e-5
        assert np.abs(sim.final_obs.cpu().numpy() - out_o["final_obs"]).max() < 2e-5
        assert np.array_equal(sim.desired.cpu().numpy(), out_o["desired"])      # goal re-sampled on reset, bit-exact
    goal, step, total, episode = sim.get_aux()
    assert (step.cpu().numpy() == 0).all() and (episode.cpu().numpy() == 2).all() and (total.cpu().numpy() == 2).all()
    qp = sim.get_state()[0].cpu().numpy()
    assert np.abs(qp[:, 8] - 0.05).max() < 1e-7                                  # cube back at the reset height
    sim.close()


def test_groups_and_graph_replay_are_bit_invariant(model_blob, monkeypatch):
    """The execution schedule (env groups on parallel streams, CUDA-graph replay, longest-first solve order, work queues
    filled through atomics, reuse of the trailing collision stage by the next step's first substep, and the slow lane that
    takes over-budget / rare-class envs through the rest of a step on its own) must not change any env's result: 4096 envs
    stepped under every combination below equal the same envs stepped as one group with plain launches and every shortcut
    off, bit for bit, including after auto-resets.  The tight-budget rows push ~half of the envs through the slow lane."""
    import torch
    from gym_so100_c_b200.engine import BatchedSim
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.rand((6, n, 6), device="cuda", generator=g) * 2 - 1
    results = []
    #          groups graph reuse slowlane budgets (newton, gjk, epa)  schedule (SO100_DAG)
    #                                                                                           K1 + K2a fused (SO100_FUSE_K12)
    configs = (("4", "1", "1", "1", None, "0", "1"), ("1", "0", "0", "0", None, "0", "0"), ("8", "0", "1", "1", None, "0", "1"),
               ("4", "1", "0", "0", None, "0", "1"), ("6", "1", "1", "1", ("1", "2", "1"), "0", "1"), ("1", "0", "1", "1", ("2", "4", "2"), "0", "0"),
               ("3", "1", "1", "0", None, "3", "1"), ("2", "0", "0", "0", None, "3", "1"), ("3", "1", "1", "0", None, "1", "0"))
    for groups, graph, reuse, slow, budgets, dag, fuse in configs:
        monkeypatch.setenv("SO100_FUSE_K12", fuse)
        monkeypatch.setenv("SO100_DAG", dag)
        monkeypatch.setenv("SO100_GROUPS", groups)
        monkeypatch.setenv("SO100_GRAPH", graph)
        monkeypatch.setenv("SO100_REUSE", reuse)
        monkeypatch.setenv("SO100_SLOWLANE", slow)
        monkeypatch.setenv("SO100_ADAPTIVE_GRIDS", "1" if budgets is None else "0")   # tight budgets: keep the slow lane on however full it gets
        for name, val in zip(("SO100_BUDGET_NEWTON", "SO100_BUDGET_GJK", "SO100_BUDGET_EPA"), budgets or ("6", "10", "5")):
            monkeypatch.setenv(name, val)
        sim = BatchedSim(n, device="cuda:0", task=0, seed=9, model_blob=model_blob)
        sim.reset()
        sim.set_aux(step_count=torch.full((n,), 697, dtype=torch.int32))     # everyone truncates (and auto-resets) at step 3
        rew = []
        for k in range(6):
            obs, r, term, trunc, succ = sim.step(acts[k], autoreset=True)
            rew.append(r.clone())
        front = 2 if fuse == "1" else 3                                            # K1 + K2a as one kernel or two
        per_stage = front + (2 if slow == "1" else (5 if dag == "1" else 3))
        assert sim.launches_per_step() == (10 * per_stage + front + 1) * int(groups)
        d = sim.diagnostics()
        results.append([t.cpu().numpy() for t in sim.get_state()] + [torch.stack(rew).cpu().numpy(), obs.cpu().numpy(),
                                                                     np.array([d["solver_runs"], d["newton_iters"], d["contacts_seen"]])])
        assert d["episodes"] >= n
        sim.close()
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [1, 3, 37])
def test_small_and_ragged_batches_match_the_oracle(model_blob, n):
    """BASELINE config 1 (a single env) and batch sizes that do not fill a block / tile group: full env.step parity."""
    import torch
    rng = np.random.default_rng(100 + n)
    sim, orc = make_pair(model_blob, n, task=0, seed=21)
    orc.reset(); sim.reset()
    a_start = np.array([0, 0.35089, -0.19493, 0, 0, -0.79585])
    for k in range(4):
        act = (a_start + rng.uniform(-0.08, 0.08, size=(n, 6))).astype(np.float32)
        out_o = orc.step(act, autoreset=False)
        obs, rew, term, trunc, succ = sim.step(torch.tensor(act), autoreset=False)
        assert obs.shape == (n, 15)
        assert np.abs(obs.cpu().numpy() - out_o["obs"]).max() < 1e-4, k
        assert np.array_equal(rew.cpu().numpy(), out_o["reward"])
    sim.close(); orc.close()


def test_full_size_batch_invariants(model_blob):
    """BASELINE config 3 at its full size (16384 envs, random actions, auto-reset): properties that do not need the oracle.
    Finite state, unit cube quaternion, joints inside their limits up to the soft-constraint give, the cube never
    tunnels through the table, rewards from the staged set, bounded contact-list overflow, and the same call repeated on
    a second handle gives bit-identical results (determinism)."""
    import torch
    from gym_so100_c_b200 import model
    from gym_so100_c_b200.engine import BatchedSim
    n = 16384
    m = model.unpack(model_blob)
    lo, hi = np.array(m["dof_range"][:6, 0]), np.array(m["dof_range"][:6, 1])
    g = torch.Generator(device="cuda").manual_seed(77)
    acts = torch.rand((25, n, 6), device="cuda", generator=g) * 2 - 1
    finals = []
    for rep in range(2):
        sim = BatchedSim(n, device="cuda:0", task=0, seed=1234, model_blob=model_blob)
        sim.reset()
        rewards = set()
        for k in range(25):
            obs, rew, term, trunc, succ = sim.step(acts[k], autoreset=True)
            rewards |= set(np.unique(rew.cpu().numpy()).tolist())
        qpos, qvel, ctrl, warm = [t.cpu().numpy() for t in sim.get_state()]
        d = sim.diagnostics()
        finals.append((qpos, qvel))
        sim.close()
        assert np.isfinite(qpos).all() and np.isfinite(qvel).all()
        assert np.abs(np.linalg.norm(qpos[:, 9:13], axis=1) - 1).max() < 1e-5
        assert (qpos[:, :6] > lo - 0.05).all() and (qpos[:, :6] < hi + 0.05).all()
        assert qpos[:, 8].min() > 0.0                         # cube centre stays above the table top (z = 0)
        assert rewards <= {0.0, 1.0, 2.0, 2.5, 3.0, 4.0}
        assert d["nonfinite_resets"] == 0 and d["contact_overflow"] <= n * 25 // 1000 and d["solver_runs"] == n * 250
        assert d["solver_cap_hits"] <= d["solver_runs"] // 100000
    assert np.array_equal(finals[0][0], finals[1][0]) and np.array_equal(finals[0][1], finals[1][1])


def test_goal_env_full_size_smoke(model_blob):
    """BASELINE config 4 shape: 65536 GoalEnv envs, dict observation pieces and the HER reward on a 4x relabel batch."""
    import torch
    from gym_so100_c_b200.engine import BatchedSim
    n = 65536
    sim = BatchedSim(n, device="cuda:0", task=1, seed=5, model_blob=model_blob)
    sim.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    for k in range(3):
        obs, rew, term, trunc, succ = sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1, autoreset=True)
    assert set(np.unique(rew.cpu().numpy()).tolist()) <= {0.0, -1.0}
    r = sim.compute_reward(sim.achieved.repeat(4, 1), sim.desired.repeat(4, 1))
    live = ~(term.bool() | trunc.bool())           # auto-reset envs already show the next episode's goals
    assert r.shape == (4 * n,) and torch.equal(r[:n][live], rew[live])      # the true goal reproduces the step reward
    assert torch.isfinite(sim.obs).all() and torch.isfinite(sim.achieved).all()
    sim.close()


def test_one_model_per_process_is_enforced(model_blob):
    """The uniform model constants live in __constant__ memory shared by all handles: a second live handle with a different
    model must be refused (loudly), the same model is fine, and after the first handle is gone a new model is accepted."""
    from gym_so100_c_b200 import ext, model
    from gym_so100_c_b200.engine import BatchedSim
    m = model.unpack(model_blob).copy()
    m["timestep"] = 0.001
    other = model.pack(m)
    a = BatchedSim(8, model_blob=model_blob)
    b = BatchedSim(8, model_blob=model_blob)
    with pytest.raises(ext.So100Error, match="one model per process"):
        BatchedSim(8, model_blob=other)
    a.close(); b.close()
    c = BatchedSim(8, model_blob=other)
    c.reset()
    c.close()
    d = BatchedSim(8, model_blob=model_blob)       # leave the process on the stock model for the tests that follow
    d.reset()
    d.close()


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


def test_seeded_reset_matches_reference_golden(model_blob):
    """VectorEnv.reset(seed=s): env i gets the reference's sample_so100_box_pose(s + i) (MT19937, utils.py:18-29)."""
    import json
    import os
    from gym_so100_c_b200.vec_env import SO100VecEnv
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))["box_pose"]
    by_seed = {g["seed"]: np.array(g["pose"]) for g in gold}
    env = SO100VecEnv(4, seed=0)
    obs, info = env.reset(seed=0)
    qp = env.get_state()[0].cpu().numpy()
    for i in range(4):
        np.testing.assert_array_equal(qp[i, 6:13], by_seed[i].astype(np.float32))
    assert obs.shape == (4, 15) and not info["is_success"].any()
    assert env.single_observation_space.contains(obs[0].cpu().numpy())
    env.close()


def test_demonstration_replay_reproduces_recorded_episodes(model_blob, tmp_path):
    """episodes recorded in the reference's pickle format (here: recorded from this env with scripted actions,
    ragged lengths) replay to the same rewards, bit for bit, with every episode in its own env."""
    import pickle
    import torch
    from gym_so100_c_b200 import replay
    from gym_so100_c_b200.vec_env import SO100VecEnv
    rng = np.random.default_rng(11)
    E, lengths = 5, [12, 7, 12, 3, 9]
    env = SO100VecEnv(E, seed=2, autoreset=False)
    obs, _ = env.reset(seed=40)
    eps = [dict(observations=[obs[i].cpu().numpy().copy()], actions=[], rewards=[], infos=[]) for i in range(E)]
    a_start = np.array([0, 0.35089, -0.19493, 0, 0, -0.79585], dtype=np.float32)
    for t in range(max(lengths)):
        act = (a_start + rng.uniform(-0.3, 0.3, size=(E, 6))).astype(np.float32)
        obs, rew, term, trunc, info = env.step(torch.from_numpy(act))
        for i in range(E):
            if t < lengths[i]:
                eps[i]["actions"].append(act[i]); eps[i]["rewards"].append(float(rew[i])); eps[i]["infos"].append({})
                eps[i]["observations"].append(obs[i].cpu().numpy().copy())
    env.close()
    path = tmp_path / "demo.pkl"
    with open(path, "wb") as f:
        pickle.dump(eps, f)
    out = replay.replay(replay.load_demonstrations(str(path)))
    assert out["reward"].shape == (12, E) and list(out["lengths"]) == lengths
    for i in range(E):
        assert np.array_equal(out["reward"][:lengths[i], i], np.array(eps[i]["rewards"], dtype=np.float32))
        assert np.abs(out["obs"][lengths[i] - 1, i] - eps[i]["observations"][-1]).max() < 1e-6
    assert np.allclose(out["episode_return"], out["recorded_return"])
    assert not out["valid"][3:, 3].any() and out["valid"][:3, 3].all()


def test_make_ids_and_sb3_terminal_observation(model_blob):
    """`make` covers the three registered ids (gym_so100/__init__.py:4-32) with their TimeLimits, and the SB3 adapter's
    terminal_observation of a GoalEnv carries the goal of the episode that ended (HerReplayBuffer stores it as next_obs)."""
    import torch
    from gym_so100_c_b200 import vec_env
    for env_id, limit in (("gym_so100/SO100CubeToBin-v0", 700), ("gym_so100/SO100TouchCube-v0", 300),
                          ("gym_so100/SO100TouchCubeSparse-v0", 300)):
        env = vec_env.make(env_id, 4, obs_type="so100_state")
        assert env.max_episode_steps == limit
        env.close()
        px = vec_env.make(env_id, 4, observation_width=32, observation_height=24)     # the registered default: pixel observations
        o, _ = px.reset()
        assert set(o) == {"pixels", "agent_pos"} and o["pixels"].shape == (4, 24, 32, 3) and o["agent_pos"].shape == (4, 6)
        px.close()
    with pytest.raises(NotImplementedError):
        vec_env.make("gym_so100/SO100Nope-v0", 4, obs_type="so100_state")
    env = vec_env.SO100GoalVecEnv(6, seed=3)
    ad = vec_env.SB3VecEnvAdapter(env)
    first = ad.reset()
    env.sim.set_aux(step_count=torch.full((6,), 299, dtype=torch.int32))
    obs, rew, done, infos = ad.step(np.zeros((6, 6), np.float32))
    assert done.all() and all(i["TimeLimit.truncated"] for i in infos)
    for i in range(6):
        term = infos[i]["terminal_observation"]
        assert np.array_equal(term["desired_goal"], first["desired_goal"][i])          # the old episode's goal
        assert not np.array_equal(obs["desired_goal"][i], first["desired_goal"][i])    # the new episode drew another one
        assert term["observation"].shape == (15,) and np.array_equal(term["achieved_goal"], term["observation"][:3])
    env.close()


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


def test_goal_vec_env_surface_and_host_step(model_blob):
    """GoalEnv dict observation, HER compute_reward (scalar and batch branches), the SB3 adapter and the
    host-buffer C-ABI entry point (so100_step_host) agree with the device path."""
    import torch
    from gym_so100_c_b200.vec_env import SB3VecEnvAdapter, SO100GoalVecEnv
    env = SO100GoalVecEnv(8, seed=2)
    obs, info = env.reset()
    assert set(obs) == {"observation", "achieved_goal", "desired_goal"} and obs["observation"].shape == (8, 15)
    act = torch.zeros((8, 6))
    obs, rew, term, trunc, info = env.step(act)
    r = env.compute_reward(obs["achieved_goal"], obs["desired_goal"], {})
    assert torch.equal(r, rew)
    assert env.compute_reward(obs["achieved_goal"][0], obs["desired_goal"][0], {}) == float(rew[0])
    assert "is_success" in info and "TimeLimit.truncated" in info
    host = env.sim.step_host(np.zeros((8, 6), np.float32))
    assert host["obs"].shape == (8, 15) and np.isfinite(host["obs"]).all() and set(np.unique(host["reward"])) <= {0.0, -1.0}
    ad = SB3VecEnvAdapter(env)
    o, r2, d, infos = ad.step(np.zeros((8, 6), np.float32))
    assert o["observation"].shape == (8, 15) and len(infos) == 8 and "is_success" in infos[0]
    assert ad.env_method("compute_reward", o["achieved_goal"], o["desired_goal"], {})[0].shape == (8,)
    env.close()


def test_host_step_pinned_pageable_and_device_paths_agree(model_blob):
    """so100_step_host copies the results out inside the step's own pipeline (per env group) when the caller's buffers are
    page-locked and after the step when they are pageable; both must return exactly what the device entry point computes,
    with and without env groups."""
    import ctypes as C
    import torch
    from gym_so100_c_b200 import ext
    from gym_so100_c_b200.engine import BatchedSim
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(11)
    acts = (torch.rand((4, n, 6), device="cuda", generator=g) * 2 - 1)
    acts_h = acts.cpu().numpy()
    sims = [BatchedSim(n, device="cuda:0", task=1, seed=21, model_blob=model_blob) for _ in range(3)]
    for s in sims:
        s.reset()
        s.set_aux(step_count=torch.full((n,), 298, dtype=torch.int32))        # truncation + auto-reset at the second step
    names = ("obs", "achieved", "desired", "reward", "terminated", "truncated", "success", "final_obs")
    page = dict(obs=np.zeros((n, 15), np.float32), achieved=np.zeros((n, 3), np.float32), desired=np.zeros((n, 3), np.float32),
                reward=np.zeros(n, np.float32), terminated=np.zeros(n, np.uint8), truncated=np.zeros(n, np.uint8),
                success=np.zeros(n, np.uint8), final_obs=np.zeros((n, 15), np.float32))
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    for k in range(4):
        dev = sims[0].step(acts[k], autoreset=True)
        pin = sims[1].step_host(acts_h[k], autoreset=True)
        ext.check(sims[2].lib.so100_step_host(sims[2].h, p(acts_h[k]), 1, p(page["obs"]), p(page["achieved"]), p(page["desired"]),
                                              p(page["reward"]), p(page["terminated"]), p(page["truncated"]), p(page["success"]),
                                              p(page["final_obs"]), sims[2]._stream()), "so100_step_host")
        ref = dict(obs=sims[0].obs, achieved=sims[0].achieved, desired=sims[0].desired, reward=sims[0].reward,
                   terminated=sims[0].terminated, truncated=sims[0].truncated, success=sims[0].success, final_obs=sims[0].final_obs)
        for name in names:
            want = ref[name].cpu().numpy()
            assert np.array_equal(pin[name], want), (k, name, "pinned")
            assert np.array_equal(page[name], want), (k, name, "pageable")
        # (envs whose curriculum goal was reached earlier restarted their step count: not everyone truncates at k = 1)
        assert (float(ref["truncated"].float().mean()) > 0.5) == (k == 1)
    for s in sims:
        s.close()


def test_contact_reuse_is_invalidated_by_state_writes(model_blob):
    """The first substep of a step reuses the contact lists the previous step left in the workspace; anything that changes the
    state between two steps (set_state, a masked reset) must switch that off for the next step.  A handle that is stepped,
    rewritten and stepped again equals a fresh handle given the same state, bit for bit."""
    import torch
    from gym_so100_c_b200.engine import BatchedSim
    n = 2048
    g = torch.Generator(device="cuda").manual_seed(3)
    acts = torch.rand((6, n, 6), device="cuda", generator=g) * 2 - 1
    a = BatchedSim(n, device="cuda:0", task=0, seed=5, model_blob=model_blob)
    b = BatchedSim(n, device="cuda:0", task=0, seed=5, model_blob=model_blob)
    a.reset(); b.reset()
    for k in range(3):
        a.step(acts[k])
    # (1) set_state: move every cube onto the far side of the table, keep the arm
    qpos, qvel, ctrl, warm = [t.clone() for t in a.get_state()]
    qpos[:, 6] = -0.05; qpos[:, 7] = 0.45; qpos[:, 8] = 0.021
    qvel[:, 6:] = 0
    a.set_state(qpos, qvel, ctrl, warm)
    b.set_state(qpos, qvel, ctrl, warm)
    goal, step, total, episode = a.get_aux()
    b.set_aux(step_count=step, total_steps=total, episode=episode)
    ra = a.step(acts[3])[1].clone(); rb = b.step(acts[3])[1].clone()
    for x, y in zip(a.get_state(), b.get_state()):
        assert torch.equal(x, y)
    assert torch.equal(ra, rb) and torch.equal(a.obs, b.obs)
    # (2) masked reset of every other env between two steps
    mask = (torch.arange(n, device="cuda") % 2).to(torch.uint8)
    a.reset(mask=mask); b.reset(mask=mask)
    ra = a.step(acts[4])[1].clone(); rb = b.step(acts[4])[1].clone()
    a.step(acts[5]); b.step(acts[5])
    for x, y in zip(a.get_state(), b.get_state()):
        assert torch.equal(x, y)
    assert torch.equal(ra, rb) and torch.equal(a.obs, b.obs)
    a.close(); b.close()


def test_unnormalize_golden_set_on_device(model_blob):
    """unnormalize_so100 (constants.py:44-47, 78-86) on the DEVICE over the whole golden action set produced by the
    reference's own Python, including the actions beyond [-1, 1] that exercise the clip: ctrl after one env.step must equal
    the reference's float32 values bit for bit."""
    import json
    import os
    import torch
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))["unnormalize_so100"]
    act = np.array(gold["action"], dtype=np.float32)
    want = np.array(gold["ctrl"], dtype=np.float32)
    assert (np.abs(act) > 1).any() and (np.abs(act) <= 1).any()          # both clipped and unclipped cases are in the set
    n = act.shape[0]
    from gym_so100_c_b200.engine import BatchedSim
    sim = BatchedSim(n, device="cuda:0", task=0, seed=1, model_blob=model_blob)
    sim.reset()
    sim.step(torch.from_numpy(act), autoreset=False)
    ctrl = sim.get_state()[2].cpu().numpy()
    assert np.array_equal(ctrl, want)
    sim.close()


def test_config3_distribution_parity(model_blob):
    """Parity on the bench's own state distribution (BASELINE config 3): 16384 envs x 60 steps of U(-1,1) actions on the GPU
    (arms on the table / base, cubes knocked about, joints at their limits), then a 1024-env subsample of that state is injected
    into the fp64 oracle and both take one more env.step with identical actions.  Per contact class of the injected state
    (no contact / box contacts only / at least one general-hull contact) the fraction of envs within
        qvel 1e-4 relative to (1 + |qvel|) after the 10 substeps, contact force 1e-3 relative to max(1, |f|) in mj_forward,
        reward / terminated / truncated / success exactly
    must be >= 99 % (no contact, box) and >= 95 % (hull).  Measured: 100 % / 100 % / 97.9-100 %."""
    import torch
    from gym_so100_c_b200.engine import BatchedSim
    from oracle.so100_oracle import Oracle
    from gym_so100_c_b200 import model
    n_big, stride, settle = 16384, 16, 60
    dev = torch.device("cuda:0")
    sim = BatchedSim(n_big, device=dev, task=0, seed=0x50100, model_blob=model_blob)
    sim.reset()
    g = torch.Generator(device=dev).manual_seed(1234)
    for s in range(settle):
        sim.step(torch.rand((n_big, 6), device=dev, generator=g) * 2 - 1, autoreset=True)
    qpos, qvel, ctrl, warm = [t[::stride].contiguous().cpu() for t in sim.get_state()]
    goal, step, total, episode = [t[::stride].contiguous().cpu() for t in sim.get_aux()]
    sim.close()
    n = qpos.shape[0]
    act = (torch.rand((n, 6), generator=torch.Generator().manual_seed(99)) * 2 - 1)
    sub = BatchedSim(n, device=dev, task=0, seed=0x50100, model_blob=model_blob)
    orc = Oracle(model_blob, n, task=0, seed=0x50100)
    sub.reset(); orc.reset()
    st64 = [t.numpy().astype(np.float64) for t in (qpos, qvel, ctrl, warm)]
    sub.set_state(qpos, qvel, ctrl, warm)
    sub.set_aux(step_count=step, total_steps=total, episode=episode)
    orc.set_state(*st64)
    orc.set_counters(step_count=step.numpy(), episode=episode.numpy().astype(np.uint32))
    # mj_forward on the injected state: contact classes and forces
    fwd = sub.forward()
    orc.forward()
    m = model.unpack(model_blob)
    hull_ids = {int(m["geom_mjid"][k]) for k in range(int(m["ngeom"])) if int(m["geom_type"][k]) == 7 and int(m["geom_vnum"][k]) != 8}
    cls, force_ok = [], []
    for i in range(n):
        oc = orc.contacts(i)
        cls.append("none" if not oc else ("hull" if any(c["geom1"] in hull_ids or c["geom2"] in hull_ids for c in oc) else "box"))
        pairs = match_contacts(gpu_contacts(fwd, i), oc)
        force_ok.append(pairs is not None and contact_errors(pairs)["force"] < 1e-3)
    cls, force_ok = np.array(cls), np.array(force_ok)
    # one env.step on both from the same state
    sub.set_state(qpos, qvel, ctrl, warm)
    orc.set_state(*st64)
    out_o = orc.step(act.numpy(), autoreset=False)
    obs, rew, term, trunc, succ = sub.step(act, autoreset=False)
    qp_g, qv_g, ctrl_g, _ = [t.cpu().numpy().astype(np.float64) for t in sub.get_state()]
    qp_o, qv_o, ctrl_o, _ = orc.get_state()
    assert np.array_equal(ctrl_g, ctrl_o)
    vel_ok = np.array([rel_err(qv_g[i], qv_o[i], floor=1.0) < 1e-4 for i in range(n)])
    flags_ok = (rew.cpu().numpy() == out_o["reward"]) & (term.cpu().numpy().astype(bool) == out_o["terminated"]) & \
               (trunc.cpu().numpy().astype(bool) == out_o["truncated"]) & (succ.cpu().numpy().astype(bool) == out_o["success"])
    ok = vel_ok & force_ok & flags_ok
    report = {c: (int((cls == c).sum()), float(ok[cls == c].mean()) if (cls == c).any() else 1.0) for c in ("none", "box", "hull")}
    print("config-3 distribution parity (envs, fraction within tolerance):", report)
    assert (cls == "box").sum() > 500 and (cls == "hull").sum() >= 20, report       # the distribution really has contacts
    assert report["none"][1] >= 0.99 and report["box"][1] >= 0.99 and report["hull"][1] >= 0.95, report
    assert flags_ok.mean() >= 0.999, float(flags_ok.mean())
    sub.close(); orc.close()


def test_episode_statistics_accumulate_on_device(model_blob):
    """RecordEpisodeStatistics equivalent (scripts/train_sac.py:290): per-env episode return / length at the step an episode
    ends, and their sums over all envs, against a host-side accumulation of the returned rewards / flags."""
    import torch
    from gym_so100_c_b200.engine import BatchedSim
    n = 512
    sim = BatchedSim(n, device="cuda:0", task=2, seed=3, model_blob=model_blob)      # shaped TouchCube reward: non-trivial returns
    sim.reset()
    sim.set_aux(step_count=torch.randint(285, 299, (n,), dtype=torch.int32))          # episodes end at different steps
    g = torch.Generator(device="cuda").manual_seed(8)
    ret = np.zeros(n); length = sim.get_aux()[1].cpu().numpy().astype(np.int64); ret_sum = 0.0; len_sum = 0; episodes = 0
    for k in range(20):
        obs, rew, term, trunc, succ = sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1, autoreset=True)
        r = rew.cpu().numpy().astype(np.float64); done = (term.bool() | trunc.bool()).cpu().numpy()
        ret += r; length += 1
        if done.any():
            np.testing.assert_allclose(sim.ep_return.cpu().numpy()[done], ret[done], rtol=1e-5, atol=1e-5)
            assert np.array_equal(sim.ep_length.cpu().numpy()[done], length[done])
            ret_sum += ret[done].sum(); len_sum += int(length[done].sum()); episodes += int(done.sum())
            ret[done] = 0; length[done] = 0
    st = sim.episode_stats()
    assert episodes >= n and st["episodes"] == episodes and st["length_sum"] == len_sum
    assert abs(st["return_sum"] - ret_sum) < 1e-3 * max(1.0, abs(ret_sum))
    sim.close()


def test_rotating_output_buffers_do_not_recapture(model_blob):
    """The step-graph cache is bounded and a caller that rotates its output buffers is moved to stable staging outputs: after
    the switch no further graph is captured, and the results equal those of a handle that always passes the same buffers."""
    import ctypes as C
    import torch
    from gym_so100_c_b200 import ext
    from gym_so100_c_b200.engine import BatchedSim
    n = 2048
    a = BatchedSim(n, device="cuda:0", task=0, seed=5, model_blob=model_blob)
    b = BatchedSim(n, device="cuda:0", task=0, seed=5, model_blob=model_blob)
    a.reset(); b.reset()
    g = torch.Generator(device="cuda").manual_seed(4)
    lib = a.lib
    p = lambda t: C.c_void_p(t.data_ptr())
    caps, hold = [], []
    for k in range(24):
        act = torch.rand((n, 6), device="cuda", generator=g) * 2 - 1
        obs = torch.empty((n, 15), device="cuda"); rew = torch.empty(n, device="cuda")       # fresh buffers every call
        term = torch.empty(n, dtype=torch.uint8, device="cuda")
        ext.check(lib.so100_step(a.h, p(act), 1, p(obs), None, None, p(rew), p(term), None, None, None, a._stream()), "so100_step")
        ref_obs, ref_rew, ref_term, _, _ = b.step(act, autoreset=True)
        assert torch.equal(obs, ref_obs) and torch.equal(rew, ref_rew) and torch.equal(term, ref_term), k
        caps.append(a.graph_stats()["captures"])
        hold.append((obs, rew, term))                                                      # keeps every call's addresses distinct
    st = a.graph_stats()
    assert st["staged"] == 1 and caps[-1] == caps[12], (st, caps)                        # no capture during the last 12 calls
    assert st["cached"] <= 4 * 9
    a.close(); b.close()


def test_her_rollout_matches_host_replay(model_blob):
    """the device-resident HER rollout (so100_her_begin / commit / sample) against a host-side replay of the same
    rollout.  Every sampled transition must be the stored one at its reported (position, env); relabelled samples must carry
    the achieved goal of a LATER step of the SAME finished episode and the reward compute_reward gives for it; 1/5 of the
    batch keeps its goal and reward (n_sampled_goal = 4, scripts/train_sac_her.py:240-244); episode statistics match."""
    import torch
    from gym_so100_c_b200.her import HerRollout
    from gym_so100_c_b200.vec_env import SO100GoalVecEnv
    from oracle.so100_oracle import compute_reward
    n, T = 256, 24
    env = SO100GoalVecEnv(n, seed=6)
    roll = HerRollout(env, horizon=T, n_sampled_goal=4, seed=3, episode_limit=8)     # episodes are cut to <= 8 steps via set_aux below
    roll.reset()
    g = torch.Generator(device="cuda").manual_seed(2)
    # episodes of 4..8 steps: start every env close to the 300-step truncation, re-arm after each reset
    env.sim.set_aux(step_count=torch.randint(292, 297, (n,), dtype=torch.int32))
    log = []
    for k in range(40):
        o = {kk: v.clone() for kk, v in env._obs().items()}
        a = torch.rand((n, 6), device="cuda", generator=g) * 2 - 1
        obs, rew, done, info = roll.step(a)
        fin = env.sim.final_obs.clone()
        log.append(dict(obs=o["observation"].cpu().numpy(), ag=o["achieved_goal"].cpu().numpy(), dg=o["desired_goal"].cpu().numpy(),
                        act=a.cpu().numpy(), rew=rew.cpu().numpy().copy(), done=done.cpu().numpy().copy(),
                        term=info["is_success"].cpu().numpy().copy(),
                        nobs=np.where(done.cpu().numpy()[:, None], fin.cpu().numpy(), obs["observation"].cpu().numpy())))
        if done.any():
            goal, step, total, ep = env.sim.get_aux()
            step = torch.where(done.cpu(), torch.randint(292, 297, (n,), dtype=torch.int32), step.cpu())
            env.sim.set_aux(step_count=step)
    # host replay of the episode bookkeeping: for each (t, env) the [first, last] step of its episode, finished or not
    first = np.zeros((40, n), int); last = np.full((40, n), -1)
    for e in range(n):
        s0 = 0
        for t in range(40):
            first[t, e] = s0
            if log[t]["done"][e]:
                last[s0:t + 1, e] = t
                s0 = t + 1
    B = 5000
    batch = {k: v.cpu().numpy() for k, v in roll.sample(B).items()}
    ix = batch["index"]
    assert (ix[:, 0] >= 0).all()
    n_real = B // 5
    assert (ix[:n_real, 2] == -1).all() and (ix[n_real:, 2] >= 0).all()
    relabel_changed = 0
    for b in range(B):
        pos, e, fut = ix[b]
        # the ring position maps to the most recent step with t % T == pos
        t = max(tt for tt in range(40) if tt % T == pos)
        assert last[t, e] >= 0 and first[t, e] > 40 - 1 - T, (b, t, e)          # finished episode, still inside the ring
        L = log[t]
        assert np.array_equal(batch["obs"][b], L["obs"][e]) and np.array_equal(batch["action"][b], L["act"][e])
        assert np.array_equal(batch["next_obs"][b], L["nobs"][e]) and np.array_equal(batch["next_achieved"][b], L["nobs"][e][:3])
        assert batch["done"][b] == (1 if L["term"][e] else 0)
        if fut < 0:
            assert np.array_equal(batch["desired"][b], L["dg"][e]) and batch["reward"][b] == L["rew"][e]
        else:
            tf = max(tt for tt in range(40) if tt % T == fut)
            assert t <= tf <= last[t, e], (b, t, tf, last[t, e])
            goal = log[tf]["nobs"][e][:3]
            assert np.array_equal(batch["desired"][b], goal)
            assert batch["reward"][b] == compute_reward(batch["next_achieved"][b][None], goal[None])[0]
            relabel_changed += int(batch["reward"][b] != L["rew"][e])
    assert relabel_changed > 0          # hindsight really turns some failures into successes (the final step always does)
    st = roll.stats()
    assert st["episodes"] == int(sum(l["done"].sum() for l in log)) and 299.0 <= st["ep_len_mean"] <= 300.0     # lengths count env steps (set_aux placed them near 300)
    env.close()


def test_pixel_observation_matches_numpy_raycaster(model_blob, model_rec):
    """obs_type "so100_pixels_agent_pos" (env.py:50-66, 130-136): the device ray-caster against the float64 numpy restatement
    of the same renderer (gym_so100_c_b200/render.py: render_numpy) on the states of a short rollout.  The two may disagree
    only at silhouette pixels (float32 vs float64 ray / facet arithmetic): >= 98.5 % of the pixels equal within 1 grey level,
    the table / cube / arm are all visible, and the red cube's pixels sit where the camera model projects the cube."""
    import torch
    from gym_so100_c_b200 import render
    from gym_so100_c_b200.vec_env import SO100VecEnv
    n, W, H = 6, 64, 48
    env = SO100VecEnv(n, obs_type="so100_pixels_agent_pos", observation_width=W, observation_height=H, seed=4)
    obs, _ = env.reset(seed=10)
    g = torch.Generator(device="cuda").manual_seed(0)
    for k in range(12):
        obs, rew, term, trunc, info = env.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
    assert obs["pixels"].dtype == torch.uint8 and obs["pixels"].shape == (n, H, W, 3)
    qpos = env.get_state()[0].cpu().numpy().astype(np.float64)
    np.testing.assert_array_equal(obs["agent_pos"].cpu().numpy(), qpos[:, :6].astype(np.float32))
    pix = obs["pixels"].cpu().numpy().astype(int)
    cube, table = int(model_rec["cg_cube"]), int(model_rec["cg_table"])
    seen_cube = 0
    for i in range(n):
        ref, hit = render.render_numpy(model_rec, qpos[i], W, H)
        same = (np.abs(pix[i] - ref.astype(int)).max(axis=-1) <= 1)
        assert same.mean() >= 0.985, (i, float(same.mean()))
        assert (hit == table).sum() > 800 and (hit >= 0).sum() > (hit == table).sum() + 30       # table, and arm / bin on top of it
        red = (pix[i][..., 0] > 100) & (pix[i][..., 1] < 30) & (pix[i][..., 2] < 30)
        assert np.array_equal(red & same, (hit == cube) & same)
        if (hit == cube).sum() >= 3:
            seen_cube += 1
            rows, cols = np.nonzero(red)
            th, asp, cz = np.tan(np.deg2rad(39.0)), W / H, 0.8 - qpos[i, 8]
            col = ((qpos[i, 6] - 0.0) / (th * asp * cz) + 1) / 2 * W - 0.5              # camera at (0, 0.6, 0.8), x right, y up
            row = (1 - (qpos[i, 7] - 0.6) / (th * cz)) / 2 * H - 0.5
            assert abs(cols.mean() - col) < 1.5 and abs(rows.mean() - row) < 1.5, (i, cols.mean(), col, rows.mean(), row)
    assert seen_cube >= 3
    env.close()


def test_goal_env_pixel_observation_layout(model_blob):
    """SO100GoalEnv's own observation layout (env.py:208-225, 267-270): flattened "top" image / 255 followed by the six joint
    angles; achieved / desired goals unchanged."""
    import torch
    from gym_so100_c_b200.vec_env import SO100GoalVecEnv
    n, W, H = 3, 32, 24
    env = SO100GoalVecEnv(n, observation="pixels", observation_width=W, observation_height=H, seed=1)
    obs, _ = env.reset()
    assert obs["observation"].shape == (n, W * H * 3 + 6) and obs["observation"].dtype == torch.float32
    assert env.single_observation_space["observation"].shape == (W * H * 3 + 6,)
    obs, rew, term, trunc, info = env.step(torch.zeros((n, 6)))
    pix = env.sim.render().reshape(n, -1).float() / 255.0
    assert torch.equal(obs["observation"][:, :-6], pix) and torch.equal(obs["observation"][:, -6:], env.sim.obs[:, 9:15])
    assert 0.0 <= float(obs["observation"][:, :-6].min()) and float(obs["observation"][:, :-6].max()) <= 1.0
    assert obs["achieved_goal"].shape == (n, 3) and set(torch.unique(rew).tolist()) <= {0.0, -1.0}
    env.close()


def test_failed_create_leaves_the_library_usable(model_blob):
    """so100_create that fails half-way (device allocation of an absurd batch) must release everything it took, including the
    live-model registration: afterwards a handle with a DIFFERENT model can be created (round 1 leaked the count and refused)."""
    import ctypes as C
    from gym_so100_c_b200 import ext, model
    from gym_so100_c_b200.engine import BatchedSim
    lib = ext.load()
    h = C.c_void_p()
    rc = lib.so100_create(model_blob, len(model_blob), 2**31 - 64, 0, 0, C.c_uint64(0), C.c_int64(0), C.byref(h))
    assert rc == -2 and b"cudaMalloc" in lib.so100_last_error()
    m = model.unpack(model_blob).copy()
    m["timestep"] = 0.001
    other = BatchedSim(8, model_blob=model.pack(m))        # would fail with "one model per process" if the failed create still counted
    other.reset()
    other.close()
    again = BatchedSim(8, model_blob=model_blob)
    again.reset()
    again.close()


def test_solver_leaves_two_point_cycles_at_the_float32_optimum(model_blob):
    """Ten states captured from the bench workload (tools/gpu_caphits.py) whose solves ran to the 100-iteration cap: the float32 iterate
    reached its optimum after a few iterations and then alternated between two points.  The stall test against the lowest cost seen must
    end them (no cap hit, a few iterations) with the same accelerations as the fp64 oracle, which converges on these states in 4-12."""
    import os
    import torch
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "solver_two_cycle_states.npz"))
    k = d["qpos"].shape[0]
    n = 64
    pad = lambda a: np.concatenate([a, np.repeat(a[:1], n - k, 0)]).astype(np.float64)
    qpos, qvel, ctrl, warm = pad(d["qpos"]), pad(d["qvel"]), pad(d["ctrl"]), pad(d["warm"])
    sim, orc = make_pair(model_blob, n)
    inject(sim, orc, qpos, qvel, ctrl, warm)
    orc.forward()
    fwd = sim.forward()
    qacc = fwd["qacc"].cpu().numpy().astype(np.float64)
    for i in range(k):
        assert rel_err(qacc[i], orc.dyn(i)["qacc"], floor=1.0) < 2e-3, (i, qacc[i], orc.dyn(i)["qacc"])
    # the same solves through the step path, where the iteration counters live
    inject(sim, orc, qpos, qvel, ctrl, warm)
    S_DIAG = 49
    before = sim.debug_read(0).view(torch.int32)[:, S_DIAG:S_DIAG + 8].clone()
    sim.substeps(1)
    after = sim.debug_read(0).view(torch.int32)[:, S_DIAG:S_DIAG + 8]
    delta = (after - before).cpu().numpy()
    assert delta[:k, 1].sum() == 0, delta[:k, 1]                  # no solve at the cap
    assert delta[:k, 5].max() <= 25, delta[:k, 5]                 # (they took 100 iterations each before)
    sim.close(); orc.close()
