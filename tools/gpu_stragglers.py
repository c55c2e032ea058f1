#!/usr/bin/env python
"""Development aid: which envs need many Newton iterations (contact sets of the stragglers)."""
import collections
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
    thr = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    m = model.load_model()
    g1, g2, mjid = m["pair_g1"], m["pair_g2"], m["geom_mjid"]
    sim = BatchedSim(n, seed=3)
    sim.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(60):
        sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
    kinds = collections.Counter()
    total = collections.Counter()
    examples = []
    for _ in range(10):
        a = sim.debug_read(0).view(torch.int32)[:, S_DIAG + 5].clone()
        sim.substeps(1)
        its = (sim.debug_read(0).view(torch.int32)[:, S_DIAG + 5] - a).cpu().numpy()
        w = sim.debug_read(1).view(torch.int32).cpu().numpy()
        qpos = sim.get_state()[0].cpu().numpy()
        for e in range(n):
            nc = min(int(w[e, 164]), 24)
            pairs = tuple(sorted((int(mjid[g1[w[e, 172 + 8 * c + 7]]]), int(mjid[g2[w[e, 172 + 8 * c + 7]]])) for c in range(nc)))
            lim = tuple(int(x) for x in np.nonzero((qpos[e, :6] < m["dof_range"][:6, 0]) | (qpos[e, :6] > m["dof_range"][:6, 1]))[0])
            key = (pairs, lim)
            total[key] += 1
            if its[e] >= thr:
                kinds[key] += 1
                if len(examples) < 12:
                    examples.append((int(its[e]), key))
    print(f"solves with >= {thr} Newton iterations, by (contact geom pairs, joints beyond their limit): count / all solves of that kind")
    for k, v in kinds.most_common(25):
        print(f"  {v:6d} / {total[k]:7d}   {k}")
    print("examples", examples)


if __name__ == "__main__":
    main()
