#!/usr/bin/env python
"""Development aid (GPU box): worst-case GPU-vs-oracle errors per scenario, to set the tolerances of tests/test_gpu_parity.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios  # noqa: E402
from gym_so100_c_b200 import model  # noqa: E402
from parity_util import contact_errors, gpu_contacts, inject, make_pair, match_contacts, rel_err  # noqa: E402

blob = model.pack(model.load_model())
N = 64
for name in scenarios.ALL:
    qpos, qvel, ctrl = scenarios.ALL[name](N)
    sim, orc = make_pair(blob, N)
    inject(sim, orc, qpos, qvel, ctrl)
    orc.forward()
    fwd = sim.forward()
    qacc_g = fwd["qacc"].cpu().numpy().astype(np.float64)
    errs = []
    for i in range(N):
        pairs = match_contacts(gpu_contacts(fwd, i), orc.contacts(i))
        e = contact_errors(pairs) if pairs is not None else dict(dist=9, pos=9, normal=9, force=9)
        e["qacc"] = rel_err(qacc_g[i], orc.dyn(i)["qacc"], floor=1.0)
        errs.append(e)
    inject(sim, orc, qpos, qvel, ctrl)
    orc.substeps(1); sim.substeps(1)
    qv_o = orc.get_state()[1]
    qv_g = sim.get_state()[1].cpu().numpy().astype(np.float64)
    ev = np.array([rel_err(qv_g[i], qv_o[i], floor=1.0) for i in range(N)])
    q = lambda k, p: float(np.percentile([e[k] for e in errs], p))
    print(f"{name:22s} " + " ".join(f"{k} max {max(e[k] for e in errs):.2e} p95 {q(k, 95):.2e}" for k in ("dist", "pos", "normal", "force", "qacc"))
          + f" qvel1 max {ev.max():.2e} p95 {np.percentile(ev, 95):.2e}")
    sim.close(); orc.close()
