#!/usr/bin/env python
"""Development aid: where the host-buffer entry point (so100_step_host) spends its time beyond the device step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = 100
sim = BatchedSim(n, seed=3)
sim.reset()
g = torch.Generator(device="cuda").manual_seed(1)
acts = torch.rand((K + 20, n, 6), device="cuda", generator=g) * 2 - 1
for s in range(20):
    sim.step(acts[s])
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(K):
    sim.step(acts[20 + s])
torch.cuda.synchronize()
back_to_back = (time.perf_counter() - t0) / K
t0 = time.perf_counter()
for s in range(K):
    sim.step(acts[20 + s])
    torch.cuda.synchronize()
sync_each = (time.perf_counter() - t0) / K
h = acts[20:].cpu().numpy()
sim.step_host(h[0])
t0 = time.perf_counter()
for s in range(K):
    sim.step_host(h[s])
host = (time.perf_counter() - t0) / K
pin = torch.empty((n, 6)).pin_memory()
t0 = time.perf_counter()
for s in range(K):
    pin.copy_(torch.from_numpy(h[s]))
    sim.step_host(pin.numpy())
host_pinned = (time.perf_counter() - t0) / K
t0 = time.perf_counter()
for s in range(K):
    sim.lib.so100_num_envs(sim.h)
ctypes_call = (time.perf_counter() - t0) / K
print(f"N={n}: device steps back to back {back_to_back * 1e3:.3f} ms; one step + sync {sync_each * 1e3:.3f} ms; step_host (pageable actions) "
      f"{host * 1e3:.3f} ms; step_host (pinned actions, incl. the host copy into them) {host_pinned * 1e3:.3f} ms; a trivial ctypes call {ctypes_call * 1e6:.1f} us")
