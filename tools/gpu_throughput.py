#!/usr/bin/env python
"""Development aid: env-steps/s of the fused step for one library build (SO100_LIB selects it)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gym_so100_c_b200.engine import BatchedSim  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    sim = BatchedSim(n, seed=3)
    sim.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand((steps + 20, n, 6), device="cuda", generator=g) * 2 - 1
    if len(sys.argv) > 3 and sys.argv[3] == "mix":
        return mix(sim, n, steps, acts)
    for s in range(int(os.environ.get("WARM", "150"))):       # settle: the bench workload's steady state needs ~150 steps
        # fresh actions every step, like bench.py (CYCLE=1: the old cycle of 20 action sets, a calmer population: numbers quoted as
        # "back to back" in DESIGN.md section 5 before item 17c's correction were taken that way)
        sim.step(acts[s % 20] if os.environ.get("CYCLE") == "1" else torch.rand((n, 6), device="cuda", generator=g) * 2 - 1)
    torch.cuda.synchronize()
    if os.environ.get("FLUSH") == "1":       # bench.py's timing: L2 flushed before every step, each step bracketed by its own event pair
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s in range(steps):
            flush.fill_(float(s))
            if os.environ.get("PREWARM") == "1":     # experiment: how much of the flushed step's extra time is the cold state / workspace
                sim.debug_read(0); sim.debug_read(1)
            ev[s][0].record()
            sim.step(acts[20 + s])
            ev[s][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            sim.step(acts[20 + s])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    d = sim.diagnostics()
    print(f"{os.path.basename(os.environ.get('SO100_LIB', 'libso100_b200.so'))} N={n}: {ms:.3f} ms/step {n / ms * 1e3:,.0f} env-steps/s "
          f"iters/solve {d['newton_iters'] / max(d['solver_runs'], 1):.2f} contacts/solve {d['contacts_seen'] / max(d['solver_runs'], 1):.2f} "
          f"overflow {d['contact_overflow']} cap {d['solver_cap_hits']} nonfinite {d['nonfinite_resets']}")


def mix(sim, n, steps, acts):
    """Half the envs random, half the pick-and-place script (bench.py's config3_mix_random_scripted)."""
    from gym_so100_c_b200 import model
    from gym_so100_c_b200.scripted import CUBE_SITE_OFFSET, ScriptedPolicy
    pol = ScriptedPolicy(model.pack(model.load_model()), n, device="cuda", period=300)
    obs = sim.obs
    pol.reset(obs[:, 0:2].double() - CUBE_SITE_OFFSET)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s in range(steps):
        a = pol.step()
        a[:n // 2] = acts[s % acts.shape[0], :n // 2]
        ev[s][0].record()
        obs, _, term, trunc, _ = sim.step(a)
        ev[s][1].record()
        pol.observe(obs, (term | trunc).bool())
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    d = sim.diagnostics()
    seg = [sum(ms[k:k + 50]) / 50 for k in range(0, steps - 49, 50)]
    print(f"mix N={n}: {sum(ms) / steps:.3f} ms/step {n * steps / sum(ms) * 1e3:,.0f} env-steps/s; per 50 steps: "
          + " ".join(f"{x:.2f}" for x in seg) + f"; successes {d['successes']} iters/solve {d['newton_iters'] / max(d['solver_runs'], 1):.2f}")


if __name__ == "__main__":
    main()
