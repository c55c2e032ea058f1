#!/usr/bin/env python
"""Print CUDA-vs-oracle error statistics per scenario (development aid; the asserting version
lives in tests/test_gpu_parity.py).  Uses the oracle as the checker only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import scenarios  # noqa: E402
from gym_so100_c_b200 import model  # noqa: E402
from parity_util import contact_errors, gpu_contacts, inject, make_pair, match_contacts, rel_err  # noqa: E402


def main():
    blob = model.pack(model.load_model())
    n = 64
    np.set_printoptions(precision=6, suppress=True, linewidth=180)
    for name, fn in scenarios.ALL.items():
        qpos, qvel, ctrl = fn(n)
        sim, orc = make_pair(blob, n)
        inject(sim, orc, qpos, qvel, ctrl)
        orc.forward()
        fwd = sim.forward()
        torch.cuda.synchronize()
        qacc_g = fwd["qacc"].cpu().numpy().astype(np.float64)
        qacc_o = np.stack([orc.dyn(i)["qacc"] for i in range(n)])
        worst = dict(dist=0.0, pos=0.0, normal=0.0, force=0.0)
        mism = 0
        for i in range(n):
            gc, oc = gpu_contacts(fwd, i), orc.contacts(i)
            pairs = match_contacts(gc, oc)
            if pairs is None:
                mism += 1
                if mism <= 2:
                    print(f"  [{name}] env {i} contact mismatch\n    gpu={[(c['geom1'], c['geom2'], round(c['dist'], 6), c['pos'].round(5)) for c in gc]}"
                          f"\n    orc={[(c['geom1'], c['geom2'], round(c['dist'], 6), c['pos'].round(5)) for c in oc]}")
                continue
            e = contact_errors(pairs)
            for k in worst:
                worst[k] = max(worst[k], e[k])
        err = np.abs(qacc_g - qacc_o) / (1 + np.abs(qacc_o))
        i, j = np.unravel_index(np.argmax(err), err.shape)
        print(f"{name:14s} forward: contact mismatches {mism}/{n}  geom err {worst}  qacc rel err max {err.max():.3e} (env {i} dof {j}: "
              f"gpu {qacc_g[i, j]:.6f} oracle {qacc_o[i, j]:.6f})  per-dof max {err.max(axis=0).round(6)}")
        orc.substeps(1)
        sim.substeps(1)
        qp_o, qv_o, _, _ = orc.get_state()
        qp_g, qv_g, _, _ = [t.cpu().numpy().astype(np.float64) for t in sim.get_state()]
        print(f"{'':14s} substep: qvel rel err {rel_err(qv_g, qv_o):.3e}  qpos abs err {np.abs(qp_g - qp_o).max():.3e}  diag {sim.diagnostics()}")
        sim.close()
    # throughput probe
    for nenv in (4096, 16384):
        sim, _ = make_pair(blob, 1)
        sim.close()
        from gym_so100_c_b200.engine import BatchedSim
        sim = BatchedSim(nenv, seed=1)
        sim.reset()
        g = torch.Generator(device="cuda").manual_seed(1234)
        act = torch.rand((nenv, 6), device="cuda", generator=g) * 2 - 1
        for _ in range(5):
            sim.step(act)
        torch.cuda.synchronize()
        t0 = time.time()
        steps = 20
        for _ in range(steps):
            act = torch.rand((nenv, 6), device="cuda", generator=g) * 2 - 1
            sim.step(act)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"throughput N={nenv}: {nenv * steps / dt:,.0f} env-steps/s  ({dt / steps * 1e3:.2f} ms/step)  diag {sim.diagnostics()}")
        sim.close()


if __name__ == "__main__":
    main()
