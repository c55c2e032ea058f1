"""Helpers that run the CUDA path (through the C ABI) and the fp64 oracle on the same injected
states and report the differences."""
from __future__ import annotations

import numpy as np


def make_pair(blob, n, task=0, seed=0, env_offset=0):
    import torch  # noqa: F401
    from gym_so100_c_b200.engine import BatchedSim
    from oracle.so100_oracle import Oracle
    sim = BatchedSim(n, device="cuda:0", task=task, seed=seed, env_offset=env_offset, model_blob=blob)
    orc = Oracle(blob, n, task=task, seed=seed, env_offset=env_offset)
    return sim, orc


def inject(sim, orc, qpos, qvel, ctrl, warm=None):
    import torch
    n = qpos.shape[0]
    warm = np.zeros((n, 12)) if warm is None else warm
    orc.set_state(qpos, qvel, ctrl, warm)
    sim.set_state(torch.tensor(qpos, dtype=torch.float32), torch.tensor(qvel, dtype=torch.float32),
                  torch.tensor(ctrl, dtype=torch.float32), torch.tensor(warm, dtype=torch.float32))


def gpu_contacts(fwd, i):
    ncon = int(fwd["ncon"][i])
    geom = fwd["con_geom"][i].cpu().numpy()
    data = fwd["con_data"][i].cpu().numpy().astype(np.float64)
    out = []
    for c in range(min(ncon, geom.shape[0])):
        out.append(dict(geom1=int(geom[c, 0]), geom2=int(geom[c, 1]), dist=data[c, 0], pos=data[c, 1:4],
                        normal=data[c, 4:7], force=data[c, 7:11]))
    return out


def _key(c):
    return (c["geom1"], c["geom2"]) + tuple(np.round(c["pos"] / 2e-4).astype(int))


def match_contacts(gc, oc):
    """Pair up contacts of the two implementations (same geom pair, nearest position)."""
    if len(gc) != len(oc):
        return None
    pairs = []
    used = set()
    for a in gc:
        best, bd = None, 1e9
        for j, b in enumerate(oc):
            if j in used or (a["geom1"], a["geom2"]) != (b["geom1"], b["geom2"]):
                continue
            d = np.linalg.norm(a["pos"] - b["pos"])
            if d < bd:
                best, bd = j, d
        if best is None:
            return None
        used.add(best)
        pairs.append((a, oc[best]))
    return pairs


def contact_errors(pairs):
    e = dict(dist=0.0, pos=0.0, normal=0.0, force=0.0)
    for a, b in pairs:
        e["dist"] = max(e["dist"], abs(a["dist"] - b["dist"]))
        e["pos"] = max(e["pos"], float(np.abs(a["pos"] - b["pos"]).max()))
        e["normal"] = max(e["normal"], float(np.abs(a["normal"] - b["normal"]).max()))
        fscale = max(1.0, float(np.abs(b["force"]).max()))
        e["force"] = max(e["force"], float(np.abs(a["force"] - b["force"]).max()) / fscale)
    return e


def rel_err(a, b, floor=1.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) / (floor + np.abs(b))).max())
