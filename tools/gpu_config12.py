#!/usr/bin/env python
"""Development aid: BASELINE configs 1 and 2 on the GPU: one env (step latency) and 4096 envs of contact-free arm motion
(actions a_start +- 0.1) through the C ABI."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

A_START = torch.tensor([0, 0.35089, -0.19493, 0, 0, -0.79585], device="cuda")


def run(n, free_space, steps=200):
    sim = BatchedSim(n, seed=3)
    sim.reset()
    if free_space:        # cube parked away from the arm (config 2: FK / dynamics / integrator only)
        qpos, qvel, ctrl, warm = sim.get_state()
        qpos[:, 6:9] = torch.tensor([0.3, 0.3, 0.02], device="cuda")
        sim.set_state(qpos, qvel, ctrl, warm)
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = (torch.rand((steps + 20, n, 6), device="cuda", generator=g) * 2 - 1) * (0.1 if free_space else 1.0) + (A_START if free_space else 0)
    for s in range(20):
        sim.step(acts[s])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        sim.step(acts[20 + s])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    d = sim.diagnostics()
    print(f"N={n} free_space={free_space}: {ms:.3f} ms/step, {n / ms * 1e3:,.0f} env-steps/s, contacts/solve {d['contacts_seen'] / max(d['solver_runs'], 1):.2f}, "
          f"iters/solve {d['newton_iters'] / max(d['solver_runs'], 1):.2f}")
    sim.close()


run(1, False)
run(4096, True)
run(4096, False)
