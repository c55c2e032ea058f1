#!/usr/bin/env python
"""Workload for ncu: 16384 envs (one env group, plain launches so that launch counts are exact), `settle` random-action
steps, then `steps` more.  With SO100_GROUPS=1 SO100_GRAPH=0 one step is exactly 64 phase_* launches:
    ncu -k regex:phase_ -s $((settle*64)) -c $((steps*64)) ... python tools/ncu_step.py settle steps"""
import os
import sys

os.environ.setdefault("SO100_GROUPS", "1")
os.environ.setdefault("SO100_GRAPH", "0")
os.environ.setdefault("SO100_FUSE_K12", "0")     # K1 and K2a as separate kernels: per-kernel attribution, 64 launches per step
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_so100_c_b200.engine import BatchedSim

settle = int(sys.argv[1]) if len(sys.argv) > 1 else 150
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
dev = torch.device("cuda:0")
sim = BatchedSim(n, device=dev, task=0, seed=0x50100)
sim.reset()
g = torch.Generator(device=dev).manual_seed(1234)
acts = torch.rand((8, n, 6), device=dev, generator=g) * 2 - 1
for s in range(settle + steps):
    sim.step(acts[s % 8], autoreset=True)
torch.cuda.synchronize()
d = sim.diagnostics()
print("launches/step", sim.launches_per_step(), "contacts/solve", d["contacts_seen"] / d["solver_runs"], "iters/solve", d["newton_iters"] / d["solver_runs"])
sim.close()
