#!/usr/bin/env python
"""Compile the reference's MJCF assets into the committed model blob + C header.

    python tools/build_model.py --assets /root/reference/gym_so100/assets

Writes gym_so100_c_b200/data/bin_a_cube.model (derived data: hull vertices, compile-time
constants) and include/so100_model.h.  The GPU box has no /root/reference, so the blob is
what the library loads by default there.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_so100_c_b200 import model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--assets", default="/root/reference/gym_so100/assets")
    args = ap.parse_args()
    m = model.load_model(args.assets)
    out = os.path.join(ROOT, "gym_so100_c_b200", "data", "bin_a_cube.model")
    with open(out, "wb") as f:
        f.write(model.pack(m))
    with open(os.path.join(ROOT, "include", "so100_model.h"), "w") as f:
        f.write(model.c_header())
    print(f"wrote {out} ({model.MODEL_DTYPE.itemsize} bytes): nbody={m['nbody']} nq={m['nq']} nv={m['nv']} "
          f"ngeom={m['ngeom']}/{m['ngeom_all']} npair={m['npair']} nvert={m['nvert']}")


if __name__ == "__main__":
    main()
