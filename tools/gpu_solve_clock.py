#!/usr/bin/env python
"""Development aid (needs a -DSO100_SOLVE_CLOCK build, SO100_LIB=...): duration vs Newton iterations of the light solves."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sim = BatchedSim(n, seed=3)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1)
for _ in range(60):
    sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
rows = []
for _ in range(20):
    sim.substeps(1)
    st = sim.debug_read(0).view(torch.int32)[:, 57:64].cpu().numpy()
    rows.append(st.copy())
    print("substep: max ns %d (its %d, ncon %d, coupled %d); mean ns %.0f; its>=8: %d" % (
        st[:, 0].max(), st[st[:, 0].argmax(), 1], st[st[:, 0].argmax(), 2] & 255, st[st[:, 0].argmax(), 2] >> 8,
        st[:, 0].mean(), (st[:, 1] >= 8).sum()))
a = np.concatenate(rows)
ok = a[:, 1] < 1000
ncon = a[:, 2] & 255
ls = a[:, 2] >> 12
ok = ok & (((a[:, 2] >> 11) & 1) == 0)     # the light kernel's solves; the dense ones are split out at the end
ok_all = a[:, 1] < 1000
for it in sorted(set(a[ok, 1].tolist())):
    m = ok & (a[:, 1] == it)
    if m.sum():
        print(f"its {it:3d}: n {m.sum():6d}  ns mean {a[m, 0].mean():9.0f}  min {a[m, 0].min():8d}  max {a[m, 0].max():8d}  ls its/newton it {ls[m].mean() / max(it, 1):5.2f}  ncon mean {ncon[m].mean():4.1f}")
print("cycle split of multi-iteration light solves (per Newton iteration): eval+driver, gradient, Hessian+factor, line search")
m = ok & (a[:, 1] >= 6)
if m.sum():
    per_it = a[m, 3:7] / a[m, 1:2]
    print(f"n {m.sum()}  cycles/iteration", np.round(per_it.mean(axis=0)).astype(int).tolist(), " ns/iteration %.0f" % (a[m, 0] / a[m, 1]).mean(),
          " ls iterations per Newton iteration %.2f (max %.1f)" % ((ls[m] / a[m, 1]).mean(), (ls[m] / a[m, 1]).max()))
m = ok & (a[:, 1] == 1)
print("single-iteration solves: cycles", np.round(a[m, 3:7].mean(axis=0)).astype(int).tolist(), " ls its %.2f" % ls[m].mean())
dense = (a[:, 2] >> 11) & 1
for name, sel in (("light kernel (block-diagonal Hessian, two envs per warp)", dense == 0), ("queue kernels (dense Hessian, one env per warp)", dense != 0)):
    m = ok_all & sel & (a[:, 1] >= 6)
    if m.sum():
        per_it = a[m, 3:7] / a[m, 1:2]
        print(f"{name}: n {m.sum()}  cycles/iteration [eval, gradient, Hessian+factor, line search] {np.round(per_it.mean(axis=0)).astype(int).tolist()}  "
              f"ns/iteration {(a[m, 0] / a[m, 1]).mean():.0f}  ls/iteration {(ls[m] / a[m, 1]).mean():.2f}  ncon {ncon[m].mean():.1f}  its mean {a[m, 1].mean():.1f} max {a[m, 1].max()}")
if os.environ.get("DENSE_SPLIT") == "1":      # a -DSO100_SOLVE_CLOCK=2 build
    m = ok_all & (dense != 0) & (a[:, 1] >= 6)
    per_it = a[m, 3:7] / a[m, 1:2]
    print("dense direction, cycles per iteration [H_c J_c rows + direction total, assembly, arm block + W, Schur + solves]", np.round(per_it.mean(axis=0)).astype(int).tolist())
