#!/usr/bin/env python
"""Development aid: env-steps/s of the fused step for one library build (SO100_LIB selects it)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    sim = BatchedSim(n, seed=3)
    sim.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand((steps + 20, n, 6), device="cuda", generator=g) * 2 - 1
    for s in range(20):
        sim.step(acts[s])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        sim.step(acts[20 + s])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    d = sim.diagnostics()
    print(f"{os.path.basename(os.environ.get('SO100_LIB', 'libso100_b200.so'))} N={n}: {ms:.3f} ms/step {n / ms * 1e3:,.0f} env-steps/s "
          f"iters/solve {d['newton_iters'] / max(d['solver_runs'], 1):.2f} contacts/solve {d['contacts_seen'] / max(d['solver_runs'], 1):.2f} "
          f"overflow {d['contact_overflow']} cap {d['solver_cap_hits']} nonfinite {d['nonfinite_resets']}")


if __name__ == "__main__":
    main()
