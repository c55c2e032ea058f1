#!/usr/bin/env python
"""Development aid: Newton iteration counts of the float32 CUDA solver vs the fp64 oracle on the same steady-state states."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200 import model  # noqa: E402
from gym_so100_c_b200.engine import BatchedSim  # noqa: E402
from oracle.so100_oracle import Oracle, build  # noqa: E402

S_DIAG = 49


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    build()
    blob = model.pack(model.load_model())
    sim = BatchedSim(n, seed=3, model_blob=blob)
    sim.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(60):
        sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
    sim.substeps(3)
    qpos, qvel, ctrl, warm = [t.cpu().numpy().astype(np.float64) for t in sim.get_state()]
    a = sim.debug_read(0).view(torch.int32)[:, S_DIAG + 5].clone()
    sim.substeps(1)
    its_g = (sim.debug_read(0).view(torch.int32)[:, S_DIAG + 5] - a).cpu().numpy()
    orc = Oracle(blob, n)
    orc.set_state(qpos, qvel, ctrl, warm)
    orc.forward()
    its_o = np.array([orc.solver(i)["iters"] for i in range(n)])
    print("gpu   iterations: mean %.3f  p99 %d  max %d" % (its_g.mean(), np.percentile(its_g, 99), its_g.max()))
    print("fp64  iterations: mean %.3f  p99 %d  max %d  (oracle tolerance 1e-11, so a few more are expected)" % (its_o.mean(), np.percentile(its_o, 99), its_o.max()))
    idx = np.argsort(-its_g)[:15]
    print("worst gpu envs: gpu its", its_g[idx].tolist(), "oracle its", its_o[idx].tolist())
    idx = np.argsort(-its_o)[:15]
    print("worst oracle envs: gpu its", its_g[idx].tolist(), "oracle its", its_o[idx].tolist())


if __name__ == "__main__":
    main()
