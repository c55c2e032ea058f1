#!/usr/bin/env python
"""Development aid: capture the pre-substep states of the solves that hit the Newton iteration cap (or need many iterations) under the
bench workload, for offline analysis with the fp64 oracle (tests/dev/analyse_caphits.py).  Writes gpurun_out/caphits.npz."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200 import model  # noqa: E402
from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

S_DIAG = 49


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    thr = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    m = model.load_model()
    lo = torch.tensor(np.asarray(m["act_lo"], dtype=np.float32).ravel()[:6], device="cuda")
    hi = torch.tensor(np.asarray(m["act_hi"], dtype=np.float32).ravel()[:6], device="cuda")
    sim = BatchedSim(n, seed=0x50100)
    sim.reset()
    g = torch.Generator(device="cuda").manual_seed(1234)
    for _ in range(150):
        sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
    rows = {k: [] for k in ("qpos", "qvel", "ctrl", "warm", "iters", "cap", "env", "substep")}
    diag = sim.debug_read(0).view(torch.int32)
    cap_prev, it_prev = diag[:, S_DIAG + 1].clone(), diag[:, S_DIAG + 5].clone()
    for s in range(steps):
        a = torch.rand((n, 6), device="cuda", generator=g) * 2 - 1
        ctrl = lo + (a + 1.0) * 0.5 * (hi - lo)
        sim.set_state(ctrl=ctrl.contiguous())
        for k in range(10):
            st = [t.clone() for t in sim.get_state()]
            sim.substeps(1)
            diag = sim.debug_read(0).view(torch.int32)
            cap, it = diag[:, S_DIAG + 1], diag[:, S_DIAG + 5]
            dc, di = cap - cap_prev, it - it_prev
            cap_prev, it_prev = cap.clone(), it.clone()
            sel = torch.nonzero((dc > 0) | (di >= thr)).flatten()
            if sel.numel():
                for name, t in zip(("qpos", "qvel", "ctrl", "warm"), st):
                    rows[name].append(t[sel].cpu().numpy())
                rows["iters"].append(di[sel].cpu().numpy()); rows["cap"].append(dc[sel].cpu().numpy())
                rows["env"].append(sel.cpu().numpy()); rows["substep"].append(np.full(sel.numel(), s * 10 + k))
    out = {k: np.concatenate(v) if v else np.zeros(0) for k, v in rows.items()}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "caphits.npz"), **out)
    print(f"{len(out['env'])} solves with >= {thr} iterations in {steps * 10} substeps of {n} envs; cap hits {int(out['cap'].sum()) if len(out['env']) else 0}")
    if len(out["env"]):
        print("iteration counts:", np.sort(out["iters"])[::-1][:40].tolist())


if __name__ == "__main__":
    main()
