#!/usr/bin/env python
"""Development aid: per-stage SM-cycle shares of the fused step (needs the -DSO100_PROFILE build:
python -c "from gym_so100_c_b200 import build; build.build(profile=True)"; SO100_LIB=.../libso100_b200_prof.so)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SO100_LIB", os.path.join(ROOT, "gym_so100_c_b200", "libso100_b200_prof.so"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200 import ext  # noqa: E402
from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

NAMES = ["kinematics+M", "bias/actuation", "collide(total)", "contact rows", "solve", "integrate", "  broad phase", "  box stage"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    mode = sys.argv[2] if len(sys.argv) > 2 else "random"
    sim = BatchedSim(n, seed=3)
    sim.reset()
    lib = ext.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    a_start = torch.tensor([0, 0.35089, -0.19493, 0, 0, -0.79585], device="cuda")
    def act():
        u = torch.rand((n, 6), device="cuda", generator=g) * 2 - 1
        return u if mode == "random" else a_start + 0.1 * u
    for _ in range(30):
        sim.step(act())
    out = (C.c_ulonglong * 8)()
    lib.so100_profile(out)
    steps = 20
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acts = [act() for _ in range(steps)]
    e0.record()
    for a in acts:
        sim.step(a)
    e1.record()
    torch.cuda.synchronize()
    lib.so100_profile(out)
    cyc = np.array(list(out), dtype=np.float64)
    tot = cyc[:6].sum()
    per_sub = cyc / (n * steps * 10)
    print(f"N={n} actions={mode}: {e0.elapsed_time(e1) / steps:.3f} ms/step (profiling build), {n * steps / (e0.elapsed_time(e1) * 1e-3):,.0f} env-steps/s")
    for k, name in enumerate(NAMES):
        extra = "" if k >= 6 else f"{100 * cyc[k] / tot:5.1f}%"
        print(f"  {name:16s} {per_sub[k]:9.0f} cycles/substep/env {extra}")
    print(f"  {'  hull stage':16s} {per_sub[2] - per_sub[6] - per_sub[7]:9.0f} cycles/substep/env")
    print("  diag", sim.diagnostics())


if __name__ == "__main__":
    main()
