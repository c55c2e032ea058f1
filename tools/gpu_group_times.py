#!/usr/bin/env python
"""Development aid: when does each env group finish within a step (SO100_GROUP_TIMES=1)?"""
import os
import sys

os.environ["SO100_GROUP_TIMES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
sim = BatchedSim(n, seed=3)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1)
acts = torch.rand((60, n, 6), device="cuda", generator=g) * 2 - 1
for s in range(50):
    sim.step(acts[s])
for s in range(50, 56):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.step(acts[s]); e1.record()
    torch.cuda.synchronize()
    print(f"step {e0.elapsed_time(e1):.3f} ms; groups done at", " ".join(f"{x:.2f}" for x in sim.group_times()))
