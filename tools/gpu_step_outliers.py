#!/usr/bin/env python
"""Development aid: per-step time of the bench workload next to the step's slowest solves (Newton iterations summed over the step's ten
substeps, per env) and largest contact lists: what the occasional slow step is made of."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

S_DIAG = 49
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
sim = BatchedSim(n, seed=0x50100)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1234)
for _ in range(150):
    sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
acts = torch.rand((steps, n, 6), device="cuda", generator=g) * 2 - 1
rows = []
prev = sim.debug_read(0).view(torch.int32)[:, S_DIAG:S_DIAG + 8].clone()
for s in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sim.step(acts[s])
    e1.record()
    torch.cuda.synchronize()
    cur = sim.debug_read(0).view(torch.int32)[:, S_DIAG:S_DIAG + 8].clone()
    d = cur - prev
    prev = cur
    w = sim.debug_read(1).view(torch.int32)
    rows.append((e0.elapsed_time(e1), int(d[:, 5].max()), int((d[:, 5] >= 100).sum()), int(d[:, 1].sum()), int(d[:, 0].sum()), int(w[:, 164].max())))
a = np.array(rows, dtype=np.float64)
print(f"{steps} steps of {n} envs: ms median {np.median(a[:, 0]):.3f} mean {a[:, 0].mean():.3f} p90 {np.percentile(a[:, 0], 90):.3f} max {a[:, 0].max():.3f}")
print("corr(ms, max per-env iterations in the step) = %.2f" % np.corrcoef(a[:, 0], a[:, 1])[0, 1])
print("slowest steps: ms, max env iterations (10 substeps), envs >= 100 iterations, cap hits, overflows, max contacts at step end")
for i in np.argsort(-a[:, 0])[:12]:
    print(f"  step {i:3d}: {a[i, 0]:.3f} ms  max-iters {int(a[i, 1]):4d}  envs>=100 {int(a[i, 2])}  cap {int(a[i, 3])}  overflow {int(a[i, 4])}  max-ncon {int(a[i, 5])}")
print("fastest steps:")
for i in np.argsort(a[:, 0])[:5]:
    print(f"  step {i:3d}: {a[i, 0]:.3f} ms  max-iters {int(a[i, 1]):4d}  envs>=100 {int(a[i, 2])}  cap {int(a[i, 3])}  overflow {int(a[i, 4])}  max-ncon {int(a[i, 5])}")
