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
ALG_BYTES_PER_ENV_STEP = 432          # 188 B read + 244 B written per env-step
FLOP_PER_ENV_STEP = 1.0e6             # SURVEY.md Appendix C convention F_contact (fp32 FLOPs, FMA = 2)
FP32_PEAK_NOMINAL_TFLOPS = 74.4       # 148 SM x 128 lanes x 2 x 1.965 GHz
METRIC = "env-steps/sec (bin-a-cube)"
# warp-instructions per env-step of the steady-state workload, from the ncu launch list in profiles/r01_phase_launches.txt
# (per 2731-env launch: solve_light 7.51 M, collide_box 3.92 M averaged over the 10 full and 1 reusing launch, kin_dyn 2.30 M,
# collide_hull 0.17 M, solve_heavy 0.05 M per launch of either instantiation, task 1.31 M):
# (10 x 7.51 + 11 x (2.30 + 3.92 + 0.17) + 20 x 0.05 + 1.31) M / 2731
WARP_INSTR_PER_ENV_STEP = 54.1e3
ISSUE_SLOTS_PER_S = 148 * 4 * 1.965e9        # SMs x schedulers x max SM clock: one warp-instruction per scheduler and cycle
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


def cpu_arm(n_envs: int, steps: int, warmup: int, seed: int = 1234):
    """The CPU implementation of the same step: the fp64 C restatement (oracle/) on all host cores.
    kind = "port": MuJoCo itself is not installable here, so this is NOT MuJoCo."""
    from gym_so100_c_b200 import model
    from oracle.so100_oracle import Oracle, build, set_threads
    build()
    cores = set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    orc = Oracle(model.pack(model.load_model()), n_envs, task=0, seed=seed)
    orc.reset()
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, size=(steps + warmup, n_envs, 6)).astype(np.float32)
    for s in range(warmup):
        orc.step(acts[s], autoreset=True)
    t0 = time.perf_counter()
    for s in range(warmup, warmup + steps):
        orc.step(acts[s], autoreset=True)
    dt = time.perf_counter() - t0
    orc.close()
    return n_envs * steps / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n = 128 * cores
    value, dt, cores = cpu_arm(n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 3: full bin-a-cube with contacts, random actions U(-1,1), auto-reset",
                   "envs_per_step": n, "note": "bounded sample of the 16384-env batch: 128 envs per host core per step"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n} envs x {args.steps} steps, fp64 C restatement of the step (oracle/), OpenMP over envs; "
                                   "MuJoCo/dm_control are not installable in this image"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


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
    sim = BatchedSim(hi - lo, device=dev, task=0, seed=0x50100, env_offset=lo)
    sim.reset()
    K, W = args.steps, args.warmup
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = torch.rand((K + W, hi - lo, 6), device=dev, generator=gen) * 2 - 1       # inputs resident in HBM
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # > 126 MB L2
    for s in range(W):
        sim.step(acts[s], autoreset=True)
    torch.cuda.synchronize()
    parallel.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall = time.perf_counter()
    for s in range(K):
        flush.fill_(float(s))                       # evict L2 between timed iterations (not timed)
        ev[s][0].record()
        sim.step(acts[W + s], autoreset=True)       # one C-ABI call = one CUDA-graph launch of the step's kernels
        ev[s][1].record()
    torch.cuda.synchronize()
    parallel.barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = parallel.max_over_ranks(float(sum(ms)), device=dev)
    kernel_ms = float(np.mean(ms))
    diag = parallel.all_reduce_stats(sim.diagnostics(), device=dev)

    # ---- per-kernel device time (CUDA events on the launching stream, inside the library) for the roofline.
    # In timing mode the library runs all envs as one group on one stream without graph replay, so that the kernels of
    # different env groups do not overlap and each duration is that kernel alone over the whole batch.
    sim.phase_timing(True)
    for s in range(min(K, 10)):
        sim.step(acts[W + s], autoreset=True)
    phase_ms, phase_cnt = sim.phase_timing(False, read=True)

    # ---- end to end through the host-buffer C-ABI call (pinned host actions -> device -> host results)
    Ke = max(3, min(K, 50))
    h_act = torch.empty((hi - lo, 6), dtype=torch.float32).pin_memory()
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
    nl = hi - lo
    h2d = nl * 6 * 4
    d2h = nl * (15 * 4 + 15 * 4 + 3 * 4 + 3 * 4 + 4 + 3)

    # ---- BASELINE config 4 beside the headline (N=1 only, short): GoalEnv dict observations for 65536 envs plus the HER
    # relabelling reward on a [65536 * 4, 3] batch every step (n_sampled_goal = 4, scripts/train_sac_her.py:242)
    extra = None
    if world == 1 and not args.no_extra:
        sim.close()
        n4 = 65536
        sim4 = BatchedSim(n4, device=dev, task=1, seed=0x50100)
        sim4.reset()
        acts4 = torch.rand((8, n4, 6), device=dev, generator=gen) * 2 - 1
        relabel = torch.rand((n4 * 4, 3), device=dev, generator=gen) * 0.1
        for s in range(5):
            sim4.step(acts4[s % 8], autoreset=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(10):
            sim4.step(acts4[s % 8], autoreset=True)
            sim4.compute_reward(sim4.achieved.repeat(4, 1), relabel)
        e1.record()
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1) / 10
        extra = {"config4_goal_env_her": {"envs": n4, "relabel_batch": n4 * 4, "ms_per_step": ms4, "value": n4 / ms4 * 1e3,
                                          "unit": "env-steps/s", "steps": 10}}
        sim4.close()
        # BASELINE config 3 with SURVEY 8d's input mix: half the envs take random actions, half run the scripted
        # pick-and-place (gym_so100_c_b200/scripted.py: reach, grasp, carry, release over the bin; episodes restart on success
        # or after 300 steps).  One whole scripted episode is timed so that every phase of it is in the figure.
        from gym_so100_c_b200 import model as _model
        from gym_so100_c_b200.scripted import CUBE_SITE_OFFSET, ScriptedPolicy
        simm = BatchedSim(hi - lo, device=dev, task=0, seed=0x50100, env_offset=lo)
        obs_m = simm.reset()[0]
        pol = ScriptedPolicy(_model.pack(_model.load_model()), hi - lo, device=dev, period=300)
        pol.reset(obs_m[:, 0:2].double() - CUBE_SITE_OFFSET)
        half = (hi - lo) // 2
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
            "envs": hi - lo, "scripted_envs": hi - lo - half, "steps": Km, "ms_per_step": msm, "value": (hi - lo) / msm * 1e3,
            "unit": "env-steps/s", "l2": "flushed between timed steps", "scripted_episodes_finished": dm["episodes"],
            "scripted_successes": dm["successes"], "contacts_per_solve": dm["contacts_seen"] / max(dm["solver_runs"], 1),
            "newton_iters_per_solve": dm["newton_iters"] / max(dm["solver_runs"], 1)}
        simm.close()
        sim = BatchedSim(hi - lo, device=dev, task=0, seed=0x50100, env_offset=lo)   # only for launches_per_step below

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        from gym_so100_c_b200.engine import measure_fp32_peak
        fp32_peak = measure_fp32_peak(local)
        n_total = n * world
        value = n_total * K / (total_ms * 1e-3)
        # dominant kernel of the step = the full-batch kernel with the largest share of device time.  The queue-driven
        # kernels (collide_hull, solve_heavy) process a few hundred items / a handful of envs per launch on a side stream;
        # their duration is the latency of their slowest item, not a share of the GPU's work (ncu launch list in profiles/).
        full_batch = [k for k in phase_ms if k not in ("collide_hull", "solve_heavy")]
        dom = max(full_batch, key=lambda k: phase_ms[k])
        dom_ms = phase_ms[dom] / max(phase_cnt[dom], 1)
        achieved = PHASE_ALG_BYTES[dom] * nl / (dom_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "phase_traffic.json")) as f:
                tj = json.load(f)
            if int(tj.get("envs", -1)) == nl:
                traffic = tj.get(dom, {}).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        fp32_tflops = FLOP_PER_ENV_STEP * nl / (kernel_ms * 1e-3) / 1e12
        cores = os.cpu_count() or 1
        cpu_n, cpu_steps = 128 * cores, 50
        # reported baseline, rank 0 at N=1 only (multi-GPU lines carry null)
        cpu_value, cpu_dt, cores = cpu_arm(cpu_n, steps=cpu_steps, warmup=2) if world == 1 else (None, None, cores)
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 3: full bin-a-cube (gripper/cube/table/bin contacts), random actions "
                                   "U(-1,1), same-step auto-reset",
                       "envs_per_gpu": n, "envs_total": n_total, "substeps_per_step": 10, "l2": "flushed between timed steps "
                       "(256 MiB write, untimed)", "parallelism": f"env-shard x{world}, no data-path collective",
                       "launches_per_step": sim.launches_per_step(), "launch": "one CUDA graph per step; env groups on parallel streams"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": f"phase_{dom}",
                         "kernel_ms_per_launch": dom_ms, "algorithmic_bytes_per_env_per_launch": PHASE_ALG_BYTES[dom],
                         "step_algorithmic_bytes_per_env": ALG_BYTES_PER_ENV_STEP,
                         "phase_share_of_step": {k: phase_ms[k] / max(sum(phase_ms.values()), 1e-9) for k in phase_ms},
                         "phase_ms_per_launch": {k: phase_ms[k] / max(phase_cnt[k], 1) for k in phase_ms},
                         "note": "HBM is not the binding roof for this path (SURVEY 8d): the step is FP32-issue bound",
                         "issue_convention": {"warp_instr_per_env_step": WARP_INSTR_PER_ENV_STEP,
                                              "achieved_warp_instr_per_s": value / world * WARP_INSTR_PER_ENV_STEP,
                                              "peak_issue_slots_per_s": ISSUE_SLOTS_PER_S,
                                              "frac": value / world * WARP_INSTR_PER_ENV_STEP / ISSUE_SLOTS_PER_S,
                                              "note": "the binding resource (DESIGN.md section 5): per-GPU share of all issue slots"},
                         "fp32_convention": {"flop_per_env_step": FLOP_PER_ENV_STEP, "achieved_tflops": fp32_tflops,
                                             "peak_tflops_nominal": FP32_PEAK_NOMINAL_TFLOPS,
                                             "peak_tflops_measured": fp32_peak,
                                             "peak_source": "FFMA loop measured in this run (so100_measure_fp32_peak)",
                                             "frac": fp32_tflops / fp32_peak,
                                             "frac_of_nominal": fp32_tflops / FP32_PEAK_NOMINAL_TFLOPS}},
            "cpu_baseline": {"value": cpu_value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{cpu_n} envs x {cpu_steps} steps of the same workload, fp64 C restatement (oracle/), OpenMP over envs"},
            "e2e": {"value": n_total * Ke / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "steps": Ke},
            "gpu_launches": K * sim.launches_per_step(),
            "clocks": clocks,
            "physics_substeps_per_s": value * 10,
            "wall_s_timed_region": t_wall,
            "diagnostics": diag,
            "extra": extra,
        }
        print(json.dumps(line))
    sim.close()
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
    ap.add_argument("--no-extra", action="store_true", help="skip the short BASELINE config 4 measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
