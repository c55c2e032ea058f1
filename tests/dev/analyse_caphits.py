#!/usr/bin/env python
"""Development aid: fp64 oracle on the states captured by tools/gpu_caphits.py (gpurun_out/caphits.npz): iteration counts, contact sets."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from gym_so100_c_b200 import model  # noqa: E402
from oracle.so100_oracle import Oracle, build  # noqa: E402


def main():
    lo = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    d = np.load(os.path.join(ROOT, "gpurun_out", "caphits.npz"))
    sel = np.nonzero(d["iters"] >= lo)[0]
    sel = sel[np.argsort(-d["iters"][sel])][:512]
    build()
    blob = model.pack(model.load_model())
    k = len(sel)
    orc = Oracle(blob, k)
    orc.set_state(d["qpos"][sel].astype(np.float64), d["qvel"][sel].astype(np.float64), d["ctrl"][sel].astype(np.float64), d["warm"][sel].astype(np.float64))
    orc.forward()
    its_o = np.array([orc.solver(i)["iters"] for i in range(k)])
    its_g = d["iters"][sel]
    print(f"{k} captured solves with >= {lo} GPU iterations: gpu mean {its_g.mean():.1f} max {its_g.max()}, oracle (1e-11) mean {its_o.mean():.1f} max {its_o.max()}")
    kinds = collections.Counter()
    for i in range(k):
        cs = orc.contacts(i)
        key = tuple(sorted((c["geom1"], c["geom2"]) for c in cs))
        kinds[key] += 1
    for key, c in kinds.most_common(12):
        print(c, key)
    for i in range(min(k, 12)):
        print("gpu", its_g[i], "oracle", orc.solver(i), "ncon", len(orc.contacts(i)))


if __name__ == "__main__":
    main()
