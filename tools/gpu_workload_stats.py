#!/usr/bin/env python
"""Development aid: per-env work distribution of the steady-state random-action workload
(contacts, box / hull candidates, GJK runs, Newton iterations per solve)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

S_DIAG = 49


def hist(name, v, bins):
    v = np.asarray(v)
    edges = list(bins) + [np.inf]
    parts = []
    for lo, hi in zip(edges[:-1], edges[1:]):
        parts.append(f"[{lo:g},{hi:g}): {100.0 * np.mean((v >= lo) & (v < hi)):.1f}%")
    print(f"{name:18s} mean {v.mean():.3f} max {v.max():.0f}  " + "  ".join(parts))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    sim = BatchedSim(n, seed=3)
    sim.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(warm):
        sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
    st0 = sim.debug_read(0).view(torch.int32)[:, S_DIAG:S_DIAG + 8].cpu().numpy().astype(np.int64)
    sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
    st1 = sim.debug_read(0).view(torch.int32)[:, S_DIAG:S_DIAG + 8].cpu().numpy().astype(np.int64)
    w = sim.debug_read(1).view(torch.int32).cpu().numpy()
    words = w.shape[1]
    stat = np.stack([w[:, 166] & 255, (w[:, 166] >> 8) & 255, (w[:, 166] >> 16) & 255, w[:, 165]], axis=1)
    ncon = w[:, 164]
    d = st1 - st0
    it = d[:, 5] / np.maximum(d[:, 6], 1)
    hist("ncon (last fwd)", ncon, [0, 1, 2, 3, 5, 9, 17, 25])
    hist("box candidates", stat[:, 0], [0, 1, 2, 4, 8, 16, 32])
    hist("box penetrating", stat[:, 1], [0, 1, 2, 3, 5, 9])
    hist("hull candidates", stat[:, 2], [0, 1, 2, 4, 8, 16, 32])
    hist("GJK runs", stat[:, 3], [0, 1, 2, 3, 5, 9])
    hist("newton it/solve", it, [0, 1, 1.05, 1.5, 2, 3, 5, 10])
    hist("contacts/solve", d[:, 7] / np.maximum(d[:, 6], 1), [0, 0.5, 1.5, 2.5, 4.5, 8.5, 16.5])
    # per-solve Newton iterations over 10 single substeps
    its = []
    for _ in range(10):
        a = sim.debug_read(0).view(torch.int32)[:, S_DIAG + 5].clone()
        sim.substeps(1)
        its.append((sim.debug_read(0).view(torch.int32)[:, S_DIAG + 5] - a).cpu().numpy())
    its = np.concatenate(its)
    hist("newton it (solve)", its, [0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 50])
    print("diag", sim.diagnostics())


if __name__ == "__main__":
    main()
