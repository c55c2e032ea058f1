#!/usr/bin/env python
"""Development aid (needs a -DSO100_TRACE build, SO100_LIB=...): per-group timeline of one step's kernels.
Prints, per env group and stage, when each kernel started / ended relative to the step's first kernel (us)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200 import ext  # noqa: E402
from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 150
sim = BatchedSim(n, seed=3)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1)
acts = torch.rand((16, n, 6), device="cuda", generator=g) * 2 - 1
for s in range(warm):
    sim.step(acts[s % 16])
lib = ext.load()
buf = np.zeros((960, 2), dtype=np.uint64)
KINDS = ["kin", "box", "hull", "lightA", "lightB", "medA", "medB", "heavy", "task", "slow"]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda") if os.environ.get("FLUSH") == "1" else None
for rep in range(2):
    lib.so100_trace_read(None)
    if flush is not None:           # bench.py's timing: the step starts with a cold L2
        flush.fill_(float(rep))
        if os.environ.get("FLUSH_SYNC", "1") == "1":
            torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sim.step(acts[rep])
    e1.record()
    torch.cuda.synchronize()
    print(f"event-timed step: {e0.elapsed_time(e1) * 1e3:.1f} us")
    lib.so100_trace_read(buf.ctypes.data_as(C.c_void_p))
    a = buf.astype(np.float64)
    valid = a[:, 1] > 0
    t0 = a[valid, 0].min()
    print(f"--- step {rep}: total {(a[valid, 1].max() - t0) / 1e3:.1f} us")
    ngroups = max(1, int(os.environ.get("SHOW_GROUPS", "2")))
    for gi in range(8):
        rows = [(st, k) for st in range(12) for k in range(10) if valid[(gi * 12 + st) * 10 + k]]
        if not rows:
            continue
        gend = max(a[(gi * 12 + st) * 10 + k, 1] for st, k in rows)
        print(f"group {gi}: ends at {(gend - t0) / 1e3:.1f} us")
        if gi >= ngroups:
            continue
        for st in range(12):
            parts = []
            for k in range(10):
                i = (gi * 12 + st) * 10 + k
                if valid[i]:
                    parts.append(f"{KINDS[k]} {(a[i, 0] - t0) / 1e3:6.1f}-{(a[i, 1] - t0) / 1e3:6.1f}")
            if parts:
                print(f"  stage {st:2d}: " + " | ".join(parts))
    # durations per kind (mean over groups and stages)
    for k in range(10):
        d = [a[(gi * 12 + st) * 10 + k, 1] - a[(gi * 12 + st) * 10 + k, 0] for gi in range(8) for st in range(12) if valid[(gi * 12 + st) * 10 + k]]
        if d:
            print(f"  {KINDS[k]:7s} n {len(d):3d}  mean {np.mean(d) / 1e3:6.1f} us  max {np.max(d) / 1e3:6.1f} us")
