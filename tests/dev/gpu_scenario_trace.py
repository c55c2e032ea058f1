#!/usr/bin/env python
"""Development aid (needs a -DSO100_SOLVE_TRACE build, SO100_LIB=...): solver trace of the env of a scenario whose contact forces are
furthest from the oracle's.  usage: gpu_scenario_trace.py <scenario>"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios  # noqa: E402
from gym_so100_c_b200 import ext, model  # noqa: E402
from parity_util import contact_errors, gpu_contacts, inject, make_pair, match_contacts, rel_err  # noqa: E402

name = sys.argv[1]
blob = model.pack(model.load_model())
N = 64
qpos, qvel, ctrl = scenarios.ALL[name](N)
sim, orc = make_pair(blob, N)
inject(sim, orc, qpos, qvel, ctrl)
orc.forward()
fwd = sim.forward()
lib = ext.load()
tr = np.zeros((64, 104, 8), dtype=np.float32)
lib.so100_solve_trace(tr.ctypes.data_as(C.c_void_p))
qacc_g = fwd["qacc"].cpu().numpy().astype(np.float64)
errs = []
for i in range(N):
    pairs = match_contacts(gpu_contacts(fwd, i), orc.contacts(i))
    e = contact_errors(pairs) if pairs is not None else dict(force=9)
    errs.append((e["force"], rel_err(qacc_g[i], orc.dyn(i)["qacc"], floor=1.0)))
errs = np.array(errs)
print("force errors sorted:", np.sort(errs[:, 0])[::-1][:8])
for i in [int(x) for x in np.argsort(-errs[:, 0])[:2]]:
    print(f"=== env {i}: force err {errs[i, 0]:.2e} qacc err {errs[i, 1]:.2e}; oracle solver {orc.solver(i)}")
    print("  it        cost      grad      gtol        gp     alpha  ls      d1  rel.pred.decrease")
    for it in range(104):
        q = tr[i, it]
        if q[0] == 0 and q[1] == 0:
            break
        print(f"  {it:3d} {q[0]:14.7f} {q[1]:9.2e} {q[2]:9.2e} {q[3]:10.3e} {q[4]:8.4f} {int(q[5]):3d} {q[6]:10.2e} {q[7]:9.2e}")
