#!/usr/bin/env python
"""bin-a-cube env-steps/s on B200 (BASELINE.json metric), plus the CPU arm.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (fp64 restatement, all host cores)

One "step" = one env.step over the whole batch: action un-normalise, 10 physics substeps,
trailing forward, reward/flags/obs, same-call auto-reset.  Workload = BASELINE
config 3: full bin-a-cube with gripper/cube/table/bin contacts, 16384 envs per GPU, random actions
U(-1,1) (the reference's action_space.sample(), scripts/example.py:20), weak scaling over GPUs.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 16384
SETTLE_STEPS = 150                    # untimed steps after reset, before --warmup: the cubes land and the arms reach the table / their
                                      # limits within ~100 steps of U(-1,1) actions (contacts per solve 0.90 -> 1.06); a window timed
                                      # right after reset is 4-5 % faster than the steady state the metric is about
ALG_BYTES_PER_ENV_STEP = 432          # 188 B read + 244 B written per env-step
FLOP_CONVENTION_PER_ENV_STEP = 1.0e6  # SURVEY.md Appendix C convention F_contact (kept as a secondary field)
FP32_PEAK_NOMINAL_TFLOPS = 74.4       # 148 SM x 128 lanes x 2 x 1.965 GHz
METRIC = "env-steps/sec (bin-a-cube)"
ISSUE_SLOTS_PER_S = 148 * 4 * 1.965e9        # SMs x schedulers x max SM clock: one warp-instruction per scheduler and cycle
CONFIG5_ENVS_PER_GPU = 131072         # BASELINE config 5: 1,048,576 envs over 8 GPUs
# Measured work of one env-step (ncu, profiles/r02_step_metrics.json: executed FP32 / FP64 thread-instructions with FFMA = 2 FLOP,
# warp-instructions, per phase kernel and per env-step, steady-state config-3 workload).  bench.py multiplies them with the
# rates it times live; the file is regenerated with tools/ncu_step.py + tools/ncu_step_summary.py whenever a kernel changes.
STEP_METRICS = os.path.join(ROOT, "profiles", "r02_step_metrics.json")
# Algorithmic bytes one env moves per launch of each phase kernel (4-byte words):
#   kin_dyn       reads qpos13+qvel12+ctrl6, writes frames102 + Marm/qfs/qas45
#   collide_box   reads frames102, writes header4 + 8 words per contact (1 contact typical)
#   collide_hull  (queued envs only, ~14 %) reads frames102 + header8, writes 8 words per contact + header
#   solve_light   reads state43 + frames102 + dyn45 + header1 + contacts8, writes state43 + counters4
#   solve_heavy   (queued envs only, < 1 % under random actions; medium <= 8 and heavy <= 24 contacts) the same per contact
#   task          reads state64 + frames102 + contacts9, writes state64 + obs15 + final_obs15 + goals6 + reward + flags
PHASE_ALG_BYTES = {"kin_dyn": (31 + 147) * 4, "collide_box": (102 + 12) * 4, "collide_hull": (110 + 9) * 4,
                   "solve_light": (199 + 47) * 4, "solve_heavy": (199 + 47) * 4, "task": (175 + 103) * 4}
# kernel launches per step: per env group 10 x (kin_dyn, collide_box, collide_hull, solve_light, solve_medium, solve_heavy) + 3 + task,
# replayed as ONE CUDA graph; the count comes from the library (so100_launches_per_step)


def step_metrics():
    try:
        with open(STEP_METRICS) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:  # noqa: BLE001
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def mujoco_arm(steps: int, warmup: int):
    """the reference env itself (MuJoCo CPU, one process per host core) when mujoco / dm_control and the
    reference package are importable.  They are not in this image, so this returns None here; the code path is
    kept so that the same bench picks the real thing up wherever it exists."""
    try:
        import mujoco  # noqa: F401
        import dm_control  # noqa: F401
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import gym_so100  # noqa: F401
    except Exception:  # noqa: BLE001
        return None
    import multiprocessing as mp
    cores = os.cpu_count() or 1

    def worker(rank, k, w, q):
        import numpy as _np
        from gym_so100.env import SO100Env
        env = SO100Env(task="so100_cube_to_bin", obs_type="so100_state")
        env._env.physics.render = lambda *a, **kw: _np.zeros((1, 1, 3), _np.uint8)      # physics only: no 3 x 640 x 480 renders
        rng = _np.random.default_rng(1234 + rank)
        env.reset(seed=rank)
        t0 = None
        for s in range(w + k):
            if s == w:
                t0 = time.perf_counter()
            _, _, term, trunc, _ = env.step(rng.uniform(-1, 1, 6).astype(_np.float32))
            if term or trunc:
                env.reset()
        q.put(time.perf_counter() - t0)

    q = mp.Queue()
    ps = [mp.Process(target=worker, args=(r, steps * 20, warmup * 20, q)) for r in range(cores)]
    for p_ in ps:
        p_.start()
    dts = [q.get() for _ in ps]
    for p_ in ps:
        p_.join()
    return cores * steps * 20 / max(dts), max(dts), cores


def cpu_arm(n_envs: int, steps: int, warmup: int, seed: int = 1234):
    """The CPU implementation of the same step on all host cores.  kind = "port": MuJoCo itself is not installable here
    (SURVEY.md 8c), so this is the repo's fp64 C restatement of the step (oracle/), built for this leg the way a production CPU
    library would be (-O3 -march=native, FMA contraction, on the host that runs it) and run with MuJoCo's own solver settings
    (tolerance 1e-8, 100 iterations, ls_tolerance 0.01; oracle_solve.c mode 1) -- not the 1e-11 checker mode the tests use."""
    from gym_so100_c_b200 import model
    from oracle import so100_oracle as O
    O.use_native()
    cores = O.set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    orc = O.Oracle(model.pack(model.load_model()), n_envs, task=0, seed=seed)
    orc.reset()
    rng = np.random.default_rng(seed)
    settle = 40                                      # the CPU arm's own settling (the cubes land within ~30 steps)
    acts = rng.uniform(-1, 1, size=(steps + warmup + settle, n_envs, 6)).astype(np.float32)
    for s in range(warmup + settle):
        orc.step(acts[s], autoreset=True)
    t0 = time.perf_counter()
    for s in range(warmup + settle, warmup + settle + steps):
        orc.step(acts[s], autoreset=True)
    dt = time.perf_counter() - t0
    orc.close()
    return n_envs * steps / dt, dt, cores


CPU_SAMPLE = ("fp64 C restatement of the step (oracle/, -O3 -march=native build, MuJoCo's solver tolerances 1e-8 / ls 0.01), "
              "OpenMP over envs; MuJoCo / dm_control are not installable in this image")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mj = mujoco_arm(args.steps, args.warmup)
    cores = os.cpu_count() or 1
    n = 64 * cores
    if mj is not None:
        value, dt, cores = mj
        kind, sample, n = "reference", f"gym_so100 SO100Env (MuJoCo CPU), one process per core, {args.steps * 20} steps each, renders patched out", cores
    else:
        value, dt, cores = cpu_arm(n, args.steps, args.warmup)
        kind, sample = "port", f"{n} envs x {args.steps} steps, " + CPU_SAMPLE
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 3: full bin-a-cube with contacts, random actions U(-1,1), auto-reset",
                   "envs_per_step": n, "note": "bounded sample of the 16384-env batch: 64 envs per host core per step, after 40 untimed "
                                               "settle steps"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def settle_steps(sim, n, steps, gen, torch, dev):
    """Untimed steps with FRESH U(-1,1) actions every step (a short cycle of repeated action sets settles into a calmer population:
    1.29 instead of 1.37 Newton iterations per solve, 7 % faster steps at 131072 envs)."""
    for _ in range(steps):
        sim.step(torch.rand((n, 6), device=dev, generator=gen) * 2 - 1, autoreset=True)


def timed_steps(sim, acts, first, count, flush, torch):
    """`count` steps from acts[first:], each bracketed by CUDA events on the launching stream, L2 flushed (untimed) before each."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count)]
    for s in range(count):
        flush.fill_(float(s))                       # evict L2 between timed iterations (not timed)
        ev[s][0].record()
        sim.step(acts[(first + s) % acts.shape[0]], autoreset=True)       # one C-ABI call = one CUDA-graph launch of the step's kernels
        ev[s][1].record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def run_gpu(args):
    import torch
    from gym_so100_c_b200 import parallel
    from gym_so100_c_b200.engine import BatchedSim

    rank, world, local = parallel.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = args.envs_per_gpu
    lo, hi = parallel.shard_range(n * world, rank, world)
    nl = hi - lo
    sim = BatchedSim(nl, device=dev, task=0, seed=0x50100, env_offset=lo)
    sim.reset()
    K, W = args.steps, args.warmup
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = torch.rand((K + W, nl, 6), device=dev, generator=gen) * 2 - 1       # inputs resident in HBM
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # > 126 MB L2
    # settle to the steady state of the workload first (always, independent of --warmup), then W warm-up steps
    settle_steps(sim, nl, args.settle, gen, torch, dev)
    for s in range(W):
        sim.step(acts[s], autoreset=True)
    torch.cuda.synchronize()
    d0 = sim.diagnostics()
    parallel.barrier()
    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("SO100_BENCH_NO_SAMPLER") != "1":
        sampler.start()
    t_wall = time.perf_counter()
    ms = timed_steps(sim, acts, W, K, flush, torch)
    parallel.barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    d1 = sim.diagnostics()
    rank_ms = parallel.gather_floats(float(sum(ms)) / K, device=dev)            # per-rank mean ms per step (rank order)
    total_ms = parallel.max_over_ranks(float(sum(ms)), device=dev)
    kernel_ms = float(np.mean(ms))
    window = {k: d1[k] - d0[k] for k in d1}
    window = parallel.all_reduce_stats(window, device=dev)
    diag = parallel.all_reduce_stats(d1, device=dev)
    ep = parallel.all_reduce_episode_stats(sim.episode_stats(), device=dev)

    # ---- per-kernel device time (CUDA events on the launching stream, inside the library) for the roofline.
    # In timing mode the library runs all envs as one group on one stream without graph replay, so that the kernels of
    # different env groups do not overlap and each duration is that kernel alone over the whole batch.
    sim.phase_timing(True)
    for s in range(min(K, 10)):
        sim.step(acts[W + s], autoreset=True)
    phase_ms, phase_cnt = sim.phase_timing(False, read=True)

    # ---- end to end through the host-buffer C-ABI call (pinned host actions -> device -> host results)
    Ke = max(3, min(K, 50))
    h_act = torch.empty((nl, 6), dtype=torch.float32).pin_memory()
    h_src = acts[W:W + Ke].cpu()
    sim.step_host(h_src[0].numpy(), autoreset=True)
    torch.cuda.synchronize()
    parallel.barrier()
    t0 = time.perf_counter()
    for s in range(Ke):
        h_act.copy_(h_src[s])
        out = sim.step_host(h_act.numpy(), autoreset=True)      # H2D + kernel + D2H + stream sync inside
        _ = float(out["reward"][0])
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0, device=dev)
    h2d = nl * 6 * 4
    d2h = nl * (15 * 4 + 15 * 4 + 3 * 4 + 3 * 4 + 4 + 3)
    launches_per_step = sim.launches_per_step()
    sim.close()

    extra = {}
    if not args.no_extra:
        # ---- BASELINE config 5 (every N): the per-GPU shard of 1,048,576 envs over 8 GPUs, 131072 envs per GPU, same workload,
        # settled, 10 timed steps; value = all ranks' envs over the slowest rank's time
        n5 = CONFIG5_ENVS_PER_GPU
        lo5, _hi5 = parallel.shard_range(n5 * world, rank, world)
        sim5 = BatchedSim(n5, device=dev, task=0, seed=0x50100, env_offset=lo5)
        sim5.reset()
        settle_steps(sim5, n5, args.settle, gen, torch, dev)
        acts5 = torch.rand((10, n5, 6), device=dev, generator=gen) * 2 - 1
        torch.cuda.synchronize()
        d50 = sim5.diagnostics()
        parallel.barrier()
        ms5 = timed_steps(sim5, acts5, 0, 10, flush, torch)
        d51 = sim5.diagnostics()
        w5 = parallel.all_reduce_stats({k: d51[k] - d50[k] for k in d51}, device=dev)
        tot5 = parallel.max_over_ranks(float(sum(ms5)), device=dev)
        rank5 = parallel.gather_floats(float(sum(ms5)) / 10, device=dev)
        sim5.close()
        extra["config5_1M_envs_over_8_gpus_shard"] = {
            "envs_per_gpu": n5, "envs_total": n5 * world, "steps": 10, "settle_steps": args.settle, "ms_per_step": tot5 / 10,
            "value": n5 * world * 10 / (tot5 * 1e-3), "unit": "env-steps/s", "rank_ms_per_step": rank5, "l2": "flushed between timed steps",
            "contacts_per_solve": w5["contacts_seen"] / max(w5["solver_runs"], 1), "newton_iters_per_solve": w5["newton_iters"] / max(w5["solver_runs"], 1)}
    if world == 1 and not args.no_extra:
        # ---- BASELINE config 4: 65536 GoalEnv envs (dict observation pieces) driving a device-resident SAC+HER rollout: every
        # step the transition goes into the HER replay ring and a batch of 65536 x n_sampled_goal = 4 relabelled goals gets its
        # reward from so100_compute_reward (scripts/train_sac_her.py:220-254)
        from gym_so100_c_b200.her import HerRollout
        from gym_so100_c_b200.vec_env import SO100GoalVecEnv
        n4 = 65536
        env4 = SO100GoalVecEnv(n4, device=dev, seed=0x50100)
        roll = HerRollout(env4, horizon=320, n_sampled_goal=4)      # ring of 320 steps x 65536 envs (3.9 GB): holds whole 300-step episodes
        roll.reset()
        # de-synchronise the episodes (every env would otherwise be truncated and reset at the same step 300, and the timed window
        # right after it would see 65536 cubes landing at once): start each env at a random point of its 300-step episode
        env4.sim.set_aux(step_count=torch.randint(0, 300, (n4,), dtype=torch.int32))
        for s in range(310):                                        # every env finishes at least one episode: the ring has finished episodes to sample
            roll.step(torch.rand((n4, 6), device=dev, generator=gen) * 2 - 1)       # fresh actions every step (see settle_steps)
        acts4 = torch.rand((10, n4, 6), device=dev, generator=gen) * 2 - 1
        roll.sample(5 * n4)                          # untimed: the first call allocates the batch tensors (0.3 - 10 ms, once)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(10):
            roll.step(acts4[s])
            batch = roll.sample(5 * n4)             # n4 real + 4 x n4 "future"-relabelled transitions (n_sampled_goal = 4), rewards recomputed
        e1.record()
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1) / 10
        extra["config4_goal_env_her"] = {"envs": n4, "sample_batch": int(batch["reward"].numel()), "relabelled": int(batch["reward"].numel()) * 4 // 5, "ms_per_step": ms4,
                                         "value": n4 / ms4 * 1e3, "unit": "env-steps/s", "steps": 10,
                                         "through": "SO100GoalVecEnv + HerRollout (device-resident ring buffer, future relabelling)",
                                         "episodes": roll.stats()}
        env4.close()
        # BASELINE config 3 with SURVEY 8d's input mix: half the envs take random actions, half run the scripted
        # pick-and-place (gym_so100_c_b200/scripted.py: reach, grasp, carry, release over the bin; episodes restart on success
        # or after 300 steps).  One whole scripted episode is timed so that every phase of it is in the figure.
        from gym_so100_c_b200 import model as _model
        from gym_so100_c_b200.scripted import CUBE_SITE_OFFSET, ScriptedPolicy
        simm = BatchedSim(nl, device=dev, task=0, seed=0x50100, env_offset=lo)
        obs_m = simm.reset()[0]
        pol = ScriptedPolicy(_model.pack(_model.load_model()), nl, device=dev, period=300)
        pol.reset(obs_m[:, 0:2].double() - CUBE_SITE_OFFSET)
        half = nl // 2
        Wm, Km = 20, 300
        evm = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Km)]
        for s in range(Wm + Km):
            a = pol.step()
            a[:half] = torch.rand((half, 6), device=dev, generator=gen) * 2 - 1
            if s >= Wm:
                flush.fill_(float(s))
                evm[s - Wm][0].record()
            obs_m, _, term_m, trunc_m, _ = simm.step(a, autoreset=True)
            if s >= Wm:
                evm[s - Wm][1].record()
            pol.observe(obs_m, (term_m | trunc_m).bool())
        torch.cuda.synchronize()
        msm = float(np.mean([a.elapsed_time(b) for a, b in evm]))
        dm = simm.diagnostics()
        extra["config3_mix_random_scripted"] = {
            "envs": nl, "scripted_envs": nl - half, "steps": Km, "ms_per_step": msm, "value": nl / msm * 1e3,
            "unit": "env-steps/s", "l2": "flushed between timed steps", "scripted_episodes_finished": dm["episodes"],
            "scripted_successes": dm["successes"], "contacts_per_solve": dm["contacts_seen"] / max(dm["solver_runs"], 1),
            "newton_iters_per_solve": dm["newton_iters"] / max(dm["solver_runs"], 1)}
        simm.close()

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        from gym_so100_c_b200.engine import measure_fp32_peak
        fp32_peak = measure_fp32_peak(local)
        n_total = n * world
        value = n_total * K / (total_ms * 1e-3)
        per_gpu_rate = nl * K / (float(sum(ms)) * 1e-3)
        # dominant kernel of the step = the full-batch kernel with the largest share of device time.  The queue-driven
        # kernels (collide_hull, solve_heavy) process a few hundred items / a handful of envs per launch on a side stream;
        # their duration is the latency of their slowest item, not a share of the GPU's work (ncu launch list in profiles/).
        full_batch = [k for k in phase_ms if k not in ("collide_hull", "solve_heavy")]
        dom = max(full_batch, key=lambda k: phase_ms[k])
        dom_ms = phase_ms[dom] / max(phase_cnt[dom], 1)
        sm_ = step_metrics()
        per = (sm_ or {}).get("per_env_step", {})
        kern = {k.split("<")[0].replace("phase_", ""): v for k, v in (sm_ or {}).get("kernels", {}).items() if "heavy" not in k}
        # measured FP32 work of the dominant kernel per env and launch (ncu thread-instruction counts, FFMA = 2)
        dom_flop = None
        if dom in kern and sm_:
            dom_flop = kern[dom]["flop32"] / kern[dom]["launches"] / sm_["envs"]
        traffic = kern[dom]["dram_bytes_per_launch"] if (dom in kern and sm_ and int(sm_["envs"]) == nl) else None
        hbm_achieved = PHASE_ALG_BYTES[dom] * nl / (dom_ms * 1e-3) / 1e9
        flop_step = per.get("flop32")
        winst_step = per.get("warp_inst")
        roof = {
            # the binding roof of this path is the FP32 (non-tensor) pipe, not HBM and not tensor cores
            "bound": "fp32", "unit": "TFLOP/s", "kernel": f"phase_{dom}", "kernel_ms_per_launch": dom_ms,
            "achieved": (dom_flop * nl / (dom_ms * 1e-3) / 1e12) if dom_flop else None,
            "peak": fp32_peak, "peak_source": "FFMA loop measured in this run (so100_measure_fp32_peak); MEASURED_PEAKS.json has no fp32 entry",
            "frac": (dom_flop * nl / (dom_ms * 1e-3) / 1e12 / fp32_peak) if dom_flop else None,
            "traffic": traffic,
            "flop32_per_env_per_launch": dom_flop, "flop_source": "ncu smsp__sass_thread_inst_executed_op_{ffma x2,fadd,fmul}_pred_on "
                                                                   "(profiles/r02_step_metrics.json)",
            "step": {"flop32_per_env_step": flop_step, "achieved_tflops": (flop_step * per_gpu_rate / 1e12) if flop_step else None,
                     "frac": (flop_step * per_gpu_rate / 1e12 / fp32_peak) if flop_step else None,
                     "peak_tflops_nominal": FP32_PEAK_NOMINAL_TFLOPS,
                     "fma_pipe_pct_by_kernel": {k: v.get("fma_pipe_pct") for k, v in kern.items()},
                     "warp_inst_per_env_step": winst_step,
                     "issue_slot_frac": (winst_step * per_gpu_rate / ISSUE_SLOTS_PER_S) if winst_step else None,
                     "note": "whole step, per GPU: measured FP32 FLOPs per env-step x env-steps/s against the measured FFMA peak; "
                             "issue_slot_frac = executed warp-instructions per second over all issue slots (the resource the step is "
                             "actually limited by, together with dependent-latency stalls)"},
            "convention_1MFLOP": {"flop_per_env_step": FLOP_CONVENTION_PER_ENV_STEP,
                                  "frac": FLOP_CONVENTION_PER_ENV_STEP * per_gpu_rate / 1e12 / fp32_peak,
                                  "note": "SURVEY Appendix C convention, superseded by the measured count"},
            "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
                    "algorithmic_bytes_per_env_per_launch": PHASE_ALG_BYTES[dom], "step_algorithmic_bytes_per_env": ALG_BYTES_PER_ENV_STEP,
                    "step_frac": ALG_BYTES_PER_ENV_STEP * per_gpu_rate / 1e9 / hbm_peak,
                    "note": "HBM is not the binding roof for this path (SURVEY 8d)"},
            "phase_share_of_step": {k: phase_ms[k] / max(sum(phase_ms.values()), 1e-9) for k in phase_ms},
            "phase_ms_per_launch": {k: phase_ms[k] / max(phase_cnt[k], 1) for k in phase_ms},
        }
        cores = os.cpu_count() or 1
        cpu_n, cpu_steps = 64 * cores, 60
        # reported baseline, rank 0 at N=1 only (multi-GPU lines carry null)
        cpu_kind = "port"
        cpu_value = None
        if world == 1:
            mj = mujoco_arm(10, 2)
            if mj is not None:
                cpu_value, _, cores = mj
                cpu_kind = "reference"
            else:
                cpu_value, _, cores = cpu_arm(cpu_n, steps=cpu_steps, warmup=2)
        solves = max(window["solver_runs"], 1)
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 3: full bin-a-cube (gripper/cube/table/bin contacts), random actions "
                                   "U(-1,1), same-step auto-reset",
                       "envs_per_gpu": n, "envs_total": n_total, "substeps_per_step": 10, "l2": "flushed between timed steps "
                       "(256 MiB write, untimed)", "parallelism": f"env-shard x{world}, no data-path collective",
                       "settle_steps": args.settle,
                       "launches_per_step": launches_per_step, "launch": "one CUDA graph per step; env groups on parallel streams"},
            "timed_window": {"contacts_per_solve": window["contacts_seen"] / solves, "newton_iters_per_solve": window["newton_iters"] / solves,
                             "solver_cap_hits": window["solver_cap_hits"], "contact_overflows": window["contact_overflow"],
                             "episodes_finished": window["episodes"]},
            "rank_ms_per_step": rank_ms,
            "step_ms": {"min": float(np.min(ms)), "median": float(np.median(ms)), "p90": float(np.percentile(ms, 90)), "max": float(np.max(ms))},
            "roofline": roof,
            "cpu_baseline": {"value": cpu_value, "unit": "env-steps/s", "cores": cores, "kind": cpu_kind,
                             "sample": f"{cpu_n} envs x {cpu_steps} steps of the same workload after 40 settle steps, " + CPU_SAMPLE},
            "e2e": {"value": n_total * Ke / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "steps": Ke},
            "gpu_launches": K * launches_per_step,
            "clocks": clocks,
            "physics_substeps_per_s": value * 10,
            "wall_s_timed_region": t_wall,
            "diagnostics": diag,
            "episode_stats": ep,
            "extra": extra or None,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--settle", type=int, default=SETTLE_STEPS, help="untimed steps after reset, before the warm-up steps")
    ap.add_argument("--no-extra", action="store_true", help="skip the BASELINE config 4 / config 5 / input-mix legs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
