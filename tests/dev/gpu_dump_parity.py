#!/usr/bin/env python
"""Development aid (GPU box): dump what the CUDA path computes on (a) the hull-contact scenarios and (b) a subsample of
the bench's own config-3 state distribution, as .npz under gpurun_out/, so that the comparison with the fp64 oracle can be
done off-line (tests/dev/analyse_parity_dump.py).  Usage: python tests/dev/gpu_dump_parity.py [settle_steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gym_so100_c_b200 import model  # noqa: E402
from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
blob = model.pack(model.load_model())


def t32(x):
    return torch.tensor(np.asarray(x), dtype=torch.float32)


def fwd_np(f):
    return {k: v.cpu().numpy() for k, v in f.items()}


def dump_scenario(name):
    z = np.load(os.path.join(ROOT, "tests", "dev", "_scen", name + ".npz"))
    n = z["qpos"].shape[0]
    sim = BatchedSim(n, device="cuda:0", task=0, seed=0, model_blob=blob)
    sim.reset()
    warm = np.zeros((n, 12))
    sim.set_state(t32(z["qpos"]), t32(z["qvel"]), t32(z["ctrl"]), t32(warm))
    f = fwd_np(sim.forward())
    sim.set_state(t32(z["qpos"]), t32(z["qvel"]), t32(z["ctrl"]), t32(warm))
    sim.substeps(1)
    st = [t.cpu().numpy() for t in sim.get_state()]
    np.savez(os.path.join(OUT, f"dump_{name}.npz"), qpos=z["qpos"], qvel=z["qvel"], ctrl=z["ctrl"], **{"fwd_" + k: v for k, v in f.items()},
             qpos1=st[0], qvel1=st[1])
    sim.close()
    print(name, "dumped", n)


def dump_config3(settle, n_big=16384, stride=16, seed=0x50100):
    dev = torch.device("cuda:0")
    sim = BatchedSim(n_big, device=dev, task=0, seed=seed, model_blob=blob)
    sim.reset()
    g = torch.Generator(device=dev).manual_seed(1234)
    for s in range(settle):
        sim.step(torch.rand((n_big, 6), device=dev, generator=g) * 2 - 1, autoreset=True)
    qpos, qvel, ctrl, warm = [t[::stride].contiguous().cpu() for t in sim.get_state()]
    goal, step, total, episode = [t[::stride].contiguous().cpu() for t in sim.get_aux()]
    d = sim.diagnostics()
    print("config3 settle", settle, "contacts/solve", d["contacts_seen"] / max(d["solver_runs"], 1), d)
    sim.close()
    n = qpos.shape[0]
    act = (torch.rand((n, 6), generator=torch.Generator().manual_seed(99)) * 2 - 1)
    sub = BatchedSim(n, device=dev, task=0, seed=seed, model_blob=blob)
    sub.reset()
    sub.set_state(qpos, qvel, ctrl, warm)
    sub.set_aux(step_count=step, total_steps=total, episode=episode)
    f = fwd_np(sub.forward())
    sub.set_state(qpos, qvel, ctrl, warm)
    sub.substeps(1)
    s1 = [t.cpu().numpy() for t in sub.get_state()]
    sub.set_state(qpos, qvel, ctrl, warm)
    obs, rew, term, trunc, succ = sub.step(act, autoreset=False)
    st = [t.cpu().numpy() for t in sub.get_state()]
    np.savez(os.path.join(OUT, f"dump_config3_s{settle}.npz"), qpos=qpos.numpy(), qvel=qvel.numpy(), ctrl=ctrl.numpy(), warm=warm.numpy(),
             step=step.numpy(), total=total.numpy(), episode=episode.numpy(), action=act.numpy(),
             **{"fwd_" + k: v for k, v in f.items()}, sub_qpos=s1[0], sub_qvel=s1[1], obs=obs.cpu().numpy(), reward=rew.cpu().numpy(),
             terminated=term.cpu().numpy(), truncated=trunc.cpu().numpy(), success=succ.cpu().numpy(),
             qpos1=st[0], qvel1=st[1], ctrl1=st[2], warm1=st[3])
    sub.close()
    print("config3 dumped", n)


if __name__ == "__main__":
    settle = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    for name in ("arm_hull_contacts", "grasp_hull_contacts"):
        dump_scenario(name)
    dump_config3(settle)
    dump_config3(settle + 240)
