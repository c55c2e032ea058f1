#!/usr/bin/env python
"""Development aid: compare the GPU dumps of tests/dev/gpu_dump_parity.py with the fp64 oracle (runs on the CPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gym_so100_c_b200 import model  # noqa: E402
from oracle.so100_oracle import Oracle  # noqa: E402
from parity_util import contact_errors, match_contacts, rel_err  # noqa: E402

m = model.load_model()
blob = model.pack(m)
HULL_IDS = {int(m["geom_mjid"][g]) for g in range(int(m["ngeom"])) if int(m["geom_type"][g]) == 7 and int(m["geom_vnum"][g]) != 8}
NAMES = {int(m["geom_mjid"][g]): g for g in range(int(m["ngeom"]))}


def gpu_contacts(z, i):
    ncon = int(z["fwd_ncon"][i])
    geom, data = z["fwd_con_geom"][i], z["fwd_con_data"][i].astype(np.float64)
    return [dict(geom1=int(geom[c, 0]), geom2=int(geom[c, 1]), dist=data[c, 0], pos=data[c, 1:4], normal=data[c, 4:7], force=data[c, 7:11])
            for c in range(min(ncon, geom.shape[0]))]


def analyse(path, warm_key=None, verbose=False):
    z = np.load(path)
    n = z["qpos"].shape[0]
    orc = Oracle(blob, n)
    warm = z[warm_key].astype(np.float64) if warm_key else np.zeros((n, 12))
    orc.set_state(z["qpos"].astype(np.float64), z["qvel"].astype(np.float64), z["ctrl"].astype(np.float64), warm)
    orc.forward()
    rows = []
    for i in range(n):
        gc, oc = gpu_contacts(z, i), orc.contacts(i)
        cls = "none" if not oc else ("hull" if any(c["geom1"] in HULL_IDS or c["geom2"] in HULL_IDS for c in oc) else "box")
        pairs = match_contacts(gc, oc)
        qa = rel_err(z["fwd_qacc"][i].astype(np.float64), orc.dyn(i)["qacc"], floor=1.0)
        if pairs is None:
            rows.append(dict(i=i, cls=cls, setdiff=True, qacc=qa, ng=len(gc), no=len(oc)))
            continue
        e = contact_errors(pairs)
        e.update(i=i, cls=cls, setdiff=False, qacc=qa, ng=len(gc), no=len(oc))
        rows.append(e)
    orc.set_state(z["qpos"].astype(np.float64), z["qvel"].astype(np.float64), z["ctrl"].astype(np.float64), warm)
    orc.substeps(1)
    qp, qv, _, _ = orc.get_state()
    key_v = "sub_qvel" if "sub_qvel" in z else "qvel1"
    key_p = "sub_qpos" if "sub_qpos" in z else "qpos1"
    for i, r in enumerate(rows):
        r["qvel1"] = rel_err(z[key_v][i].astype(np.float64), qv[i], floor=1.0)
        r["qpos1"] = float(np.abs(z[key_p][i].astype(np.float64) - qp[i]).max())
    print(os.path.basename(path), "n", n)
    for cls in ("none", "box", "hull"):
        R = [r for r in rows if r["cls"] == cls]
        if not R:
            continue
        sd = sum(r["setdiff"] for r in R)
        ok = [r for r in R if not r["setdiff"]]
        f = lambda k, tol: np.mean([r.get(k, 0) < tol for r in ok]) if ok else float("nan")
        print(f"  {cls:5s} {len(R):5d} envs, contact-set differs {sd}; within: pos 2e-5 {f('pos', 2e-5):.3f}  normal 2e-4 {f('normal', 2e-4):.3f}  dist 2e-6 {f('dist', 2e-6):.3f}"
              f"  force 1e-3 {f('force', 1e-3):.3f} 5e-3 {f('force', 5e-3):.3f}  qacc 2e-3 {f('qacc', 2e-3):.3f} 1e-4 {f('qacc', 1e-4):.3f}"
              f"  qvel1 1e-4 {np.mean([r['qvel1'] < 1e-4 for r in R]):.3f} 1e-5 {np.mean([r['qvel1'] < 1e-5 for r in R]):.3f}")
    if verbose:
        for r in rows:
            if r["setdiff"] or r.get("pos", 0) > 2e-5 or r["qvel1"] > 1e-4:
                print("   ", {k: (float(f"{v:.3g}") if isinstance(v, float) else v) for k, v in r.items()})
    orc.close()
    return rows, z


if __name__ == "__main__":
    for p in sys.argv[1:]:
        analyse(p, warm_key="warm" if "config3" in p else None, verbose=os.environ.get("V") == "1")
