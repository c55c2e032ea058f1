#!/usr/bin/env python
"""Development aid: per-kernel-class device time of the step pipeline (steady-state random actions)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    sim = BatchedSim(n, seed=3)
    sim.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand((steps + 40, n, 6), device="cuda", generator=g) * 2 - 1
    for s in range(40):
        sim.step(acts[s])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        sim.step(acts[40 + s])
    e1.record()
    torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) / steps
    sim.phase_timing(True)
    for s in range(steps):
        sim.step(acts[40 + s])
    ms, cnt = sim.phase_timing(False, read=True)
    tot = sum(ms.values())
    print(f"{os.path.basename(os.environ.get('SO100_LIB', 'libso100_b200.so'))} N={n}: {plain:.3f} ms/step = {n / plain * 1e3:,.0f} env-steps/s; "
          f"sum of kernels {tot / steps:.3f} ms/step")
    for k in ms:
        print(f"  {k:14s} {1e3 * ms[k] / max(cnt[k], 1):8.1f} us/launch x {cnt[k] // steps:3d}  {100 * ms[k] / tot:5.1f}%")


if __name__ == "__main__":
    main()
