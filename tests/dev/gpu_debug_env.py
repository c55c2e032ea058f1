#!/usr/bin/env python
"""Development aid: dump contacts / qacc of single envs of a scenario for the CUDA path and the oracle.
usage: gpu_debug_env.py <scenario> <env> [<env> ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenarios  # noqa: E402
from gym_so100_c_b200 import model  # noqa: E402
from parity_util import gpu_contacts, inject, make_pair  # noqa: E402


def main():
    name = sys.argv[1]
    envs = [int(a) for a in sys.argv[2:]]
    np.set_printoptions(precision=5, suppress=True, linewidth=200)
    blob = model.pack(model.load_model())
    n = 64
    qpos, qvel, ctrl = scenarios.ALL[name](n)
    sim, orc = make_pair(blob, n)
    inject(sim, orc, qpos, qvel, ctrl)
    orc.forward()
    fwd = sim.forward()
    qacc_g = fwd["qacc"].cpu().numpy()
    for i in envs:
        print(f"=== {name} env {i}: oracle solver {orc.solver(i)}")
        print("qpos", qpos[i]); print("qvel", qvel[i])
        print("qacc gpu   ", qacc_g[i])
        print("qacc oracle", orc.dyn(i)["qacc"])
        print("qacc_smooth oracle", orc.dyn(i)["qacc_smooth"])
        for tag, cs in (("gpu", gpu_contacts(fwd, i)), ("orc", orc.contacts(i))):
            for c in cs:
                print(f"  {tag} ({c['geom1']:2d},{c['geom2']:2d}) dist {c['dist']:.6f} pos {c['pos'].round(5)} n {c['normal'].round(4)} f {np.asarray(c['force']).round(4)}")
    print(sim.diagnostics())


if __name__ == "__main__":
    main()
