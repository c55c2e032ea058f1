#!/usr/bin/env python
"""Development aid (needs a -DSO100_SOLVE_TRACE build, SO100_LIB=...): per-iteration trace of the float32 Newton solver on the states
captured by tools/gpu_caphits.py, next to the fp64 oracle's iteration count."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200 import ext  # noqa: E402
from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

src = os.path.join(ROOT, "gpurun_out", "caphits.npz")
d = np.load(src if os.path.exists(src) else os.path.join(ROOT, "tests", "dev", "_scen", "caphits.npz"))
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 30
hi = int(os.environ.get("HI", "1000"))
sel = np.nonzero((d["iters"] >= lo) & (d["iters"] <= hi))[0]
sel = sel[np.argsort(-d["iters"][sel])][:64]
k = len(sel)
pad = lambda a: np.concatenate([a[sel], np.repeat(a[sel][:1], 64 - k, 0)]).astype(np.float32)
sim = BatchedSim(64, seed=1)
sim.reset()
sim.set_state(*[torch.tensor(pad(d[n]), device="cuda") for n in ("qpos", "qvel", "ctrl", "warm")])
sim.forward()
lib = ext.load()
tr = np.zeros((64, 104, 8), dtype=np.float32)
lib.so100_solve_trace(tr.ctypes.data_as(C.c_void_p))
np.save(os.path.join(ROOT, "gpurun_out", "solve_trace.npy"), tr)
np.set_printoptions(linewidth=200)
show = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for i in range(min(k, show)):
    print(f"=== capture {sel[i]}: env {d['env'][sel[i]]} substep {d['substep'][sel[i]]}, {d['iters'][sel[i]]} iterations in the step")
    print("  it        cost      grad      gtol        gp     alpha  ls      d1  rel.pred.decrease")
    for it in range(104):
        q = tr[i, it]
        if q[0] == 0 and q[1] == 0:
            break
        if it < 25 or it % 10 == 0:
            print(f"  {it:3d} {q[0]:14.7f} {q[1]:9.2e} {q[2]:9.2e} {q[3]:10.3e} {q[4]:8.4f} {int(q[5]):3d} {q[6]:10.2e} {q[7]:9.2e}")
