#!/usr/bin/env python
"""Summarise the per-launch ncu metric list of tools/ncu_step.py (one env group, 64 launches per step): per phase kernel
the device time, executed FP32 / FP64 thread-instructions (FFMA counted as 2 FLOP), FMA-pipe utilisation, issue-slot
utilisation, active lanes per instruction and DRAM bytes; then the per-env-step totals bench.py's roofline uses.
    python tools/ncu_step_summary.py gpurun_out/r02_step_metrics.csv ENVS STEPS [out.json]"""
import collections
import csv
import json
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].replace("void ", "").replace("so100::", "")
        name = name.split("(")[0]
        per.setdefault((int(r[iid]), name), {})[r[im]] = float(r[iv].replace(",", ""))
    return per


def main(path, envs, steps, out=None):
    per = load(path)
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for (_, k), m in per.items():
        for name, v in m.items():
            agg[k][name].append(v)
    S = lambda v, n: sum(v.get(n, [0.0]))
    M = lambda v, n: (sum(v[n]) / len(v[n])) if n in v else float("nan")
    tot_t = sum(S(v, "gpu__time_duration.sum") for v in agg.values())
    print(f"{'kernel':34s} {'n':>3s} {'ms':>7s} {'share':>6s} {'us/l':>7s} {'GFLOP32':>8s} {'GFLOP64':>8s} {'fma%':>6s} {'issue%':>6s} {'lanes':>5s} {'warps%':>6s} {'Mwinst':>7s} {'DRAM MB/l':>9s}")
    summ = {}
    tot = collections.defaultdict(float)
    for k, v in sorted(agg.items(), key=lambda kv: -S(kv[1], "gpu__time_duration.sum")):
        t = S(v, "gpu__time_duration.sum")
        n = len(v["gpu__time_duration.sum"])
        f32 = 2 * S(v, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum") + S(v, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum") + \
            S(v, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum")
        f64 = 2 * S(v, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum") + S(v, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum") + \
            S(v, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
        winst = S(v, "smsp__inst_executed.sum")
        dram = S(v, "dram__bytes_read.sum") + S(v, "dram__bytes_write.sum")
        # time-weighted means of the utilisation metrics
        tw = lambda name: (sum(a * b for a, b in zip(v[name], v["gpu__time_duration.sum"])) / t) if name in v and t else float("nan")
        row = dict(launches=n, ms=t / 1e6, share=t / tot_t, us_per_launch=t / n / 1e3, flop32=f32, flop64=f64,
                   fma_pipe_pct=tw("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                   issue_pct=tw("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   lanes_per_inst=S(v, "smsp__thread_inst_executed.sum") / max(winst, 1),
                   warps_pct=tw("sm__warps_active.avg.pct_of_peak_sustained_active"), warp_inst=winst, dram_bytes_per_launch=dram / n,
                   fma_pipe_inst=S(v, "sm__inst_executed_pipe_fma.sum"), alu_pipe_inst=S(v, "sm__inst_executed_pipe_alu.sum"),
                   fp64_pipe_inst=S(v, "sm__inst_executed_pipe_fp64.sum"), lsu_pipe_inst=S(v, "sm__inst_executed_pipe_lsu.sum"),
                   xu_pipe_inst=S(v, "sm__inst_executed_pipe_xu.sum"))
        summ[k] = row
        for key in ("flop32", "flop64", "warp_inst", "fma_pipe_inst", "alu_pipe_inst", "fp64_pipe_inst", "lsu_pipe_inst", "xu_pipe_inst"):
            tot[key] += row[key]
        tot["dram"] += dram
        print(f"{k[:34]:34s} {n:3d} {t / 1e6:7.3f} {100 * t / tot_t:5.1f}% {t / n / 1e3:7.1f} {f32 / 1e9:8.3f} {f64 / 1e9:8.3f} "
              f"{row['fma_pipe_pct']:6.1f} {row['issue_pct']:6.1f} {row['lanes_per_inst']:5.1f} {row['warps_pct']:6.1f} {winst / 1e6:7.2f} {dram / n / 1e6:9.2f}")
    es = envs * steps
    per_env_step = {k: v / es for k, v in tot.items()}
    per_env_step["serialised_us_per_step"] = tot_t / steps / 1e3
    print(f"sum {tot_t / 1e6:.3f} ms over {sum(r['launches'] for r in summ.values())} launches, {steps} steps of {envs} envs")
    print("per env-step:", {k: round(v, 1) for k, v in per_env_step.items()})
    if out:
        json.dump({"envs": envs, "steps": steps, "kernels": summ, "per_env_step": per_env_step}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else None)
