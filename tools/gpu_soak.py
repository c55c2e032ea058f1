#!/usr/bin/env python
"""Development aid: long random-action rollout (truncations, auto-resets) and its diagnostics."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
task = int(sys.argv[3]) if len(sys.argv) > 3 else 0
sim = BatchedSim(n, seed=11, task=task)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(2)
t0 = time.perf_counter()
rsum = torch.zeros(n, device="cuda")
rmax = torch.full((n,), -10.0, device="cuda")
for s in range(steps):
    obs, rew, term, trunc, succ = sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
    rsum += rew
    rmax = torch.maximum(rmax, rew)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
d = sim.diagnostics()
qpos, qvel, _, _ = sim.get_state()
print(f"task {task}: {n} envs x {steps} steps in {dt:.2f} s wall ({n * steps / dt / 1e6:.2f} M env-steps/s incl. action sampling)")
print("diagnostics", d)
print("per solve: newton iters %.3f, contacts %.3f; cap hits %.2e of solves; overflow %.2e of env-steps; nonfinite resets %d" % (
    d["newton_iters"] / d["solver_runs"], d["contacts_seen"] / d["solver_runs"], d["solver_cap_hits"] / d["solver_runs"],
    d["contact_overflow"] / (n * steps), d["nonfinite_resets"]))
print("episodes finished %d (expected about %d from truncation alone), successes %d" % (
    d["episodes"], n * (steps // (700 if task == 0 else 300)), d["successes"]))
print("state: finite %s, |qvel| max %.2f, cube z min %.4f max %.3f, reward max over run: %s" % (
    bool(torch.isfinite(qpos).all() and torch.isfinite(qvel).all()), float(qvel.abs().max()), float(qpos[:, 8].min()),
    float(qpos[:, 8].max()), sorted(set(rmax.cpu().numpy().round(2).tolist()))[-4:]))
