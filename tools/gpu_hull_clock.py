#!/usr/bin/env python
"""Development aid (needs a -DSO100_HULL_CLOCK build, SO100_LIB=...): latency of the GJK/EPA items."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gym_so100_c_b200 import ext  # noqa: E402
from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sim = BatchedSim(n, seed=3)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1)
for _ in range(60):
    sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
lib = ext.load()
buf = np.zeros((65536, 4), dtype=np.int32)
lib.so100_hull_stats(buf.ctypes.data_as(C.c_void_p))
for _ in range(3):
    sim.step(torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
ph = np.zeros((65536, 6), dtype=np.int32)
lib.so100_hull_phases(ph.ctypes.data_as(C.c_void_p))
cnt = lib.so100_hull_stats(buf.ctypes.data_as(C.c_void_p))
a = buf[:min(cnt, 65536)]
gjk_ns, post_ns = a[:, 0], a[:, 1]
gi, ei, hit = a[:, 2] & 255, (a[:, 2] >> 8) & 255, (a[:, 2] >> 16) & 1
print(f"{len(a)} items over 33 position stages; hits {hit.mean() * 100:.1f}%")
print("separated: gjk ns mean %.0f p99 %.0f max %d; gjk its mean %.1f max %d" % (
    gjk_ns[hit == 0].mean(), np.percentile(gjk_ns[hit == 0], 99), gjk_ns[hit == 0].max(), gi[hit == 0].mean(), gi[hit == 0].max()))
h = hit == 1
print("hits: gjk+epa ns mean %.0f p90 %.0f p99 %.0f max %d; gjk its mean %.1f; epa its mean %.1f p90 %.0f max %d; post ns mean %.0f max %d" % (
    gjk_ns[h].mean(), np.percentile(gjk_ns[h], 90), np.percentile(gjk_ns[h], 99), gjk_ns[h].max(), gi[h].mean(), ei[h].mean(),
    np.percentile(ei[h], 90), ei[h].max(), post_ns[h].mean(), post_ns[h].max()))
for e in sorted(set(ei[h].tolist())):
    m = h & (ei == e)
    print(f"  epa its {e:2d}: n {m.sum():5d}  ns mean {gjk_ns[m].mean():8.0f}  verts mean {a[m, 3].mean():5.0f}")

ph = ph[:len(a)]
for lo, hi in ((1, 4), (5, 8), (9, 30)):
    m = h & (ei >= lo) & (ei <= hi)
    if m.sum():
        per = ph[m].sum(0) / ei[m].sum()
        print(f"  epa its {lo}-{hi}: cycles per iteration: closest {per[0]:.0f} support {per[1]:.0f} visibility {per[2]:.0f} horizon {per[3]:.0f} new faces {per[4]:.0f}  (verts mean {a[m, 3].mean():.0f})")
