"""bench.py contract checks that run without a GPU: the CPU arm (`--impl reference`) prints one JSON line with the required
keys, and the static pieces of the GPU arm's line (metric name, algorithmic-bytes table, launch accounting) are consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)   # not the OMP_NUM_THREADS=1 of torchrun
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["value"] > 0 and "workload" in d["config"]


def test_static_tables_of_the_gpu_arm():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.METRIC.startswith("env-steps/sec")
    assert set(bench.PHASE_ALG_BYTES) == {"kin_dyn", "collide_box", "collide_hull", "solve_light", "solve_heavy", "task"}
    assert bench.ALG_BYTES_PER_ENV_STEP == 432                      # 188 B read + 244 B written
    traffic = json.load(open(os.path.join(ROOT, "profiles", "phase_traffic.json")))
    # measured DRAM traffic of the dominant kernel within 2x of its algorithmic bytes (no wasted re-reads)
    alg = bench.PHASE_ALG_BYTES["solve_light"] * traffic["envs"]
    assert 0.5 * alg < traffic["solve_light"]["dram_bytes_per_launch"] < 2.0 * alg


def test_measured_step_metrics_feed_the_roofline():
    """The FP32 roofline numerator is MEASURED (ncu thread-instruction counts, profiles/r02_step_metrics.json), not the
    Appendix-C convention: the file bench.py reads must exist, cover every phase kernel and be self-consistent."""
    sys.path.insert(0, ROOT)
    import bench
    m = bench.step_metrics()
    assert m is not None and m["envs"] == bench.ENVS_PER_GPU and m["steps"] >= 1
    kernels = {k.split("<")[0].replace("phase_", "") for k in m["kernels"]}
    assert {"kin_dyn", "collide_box", "collide_hull", "solve_light", "solve_heavy", "task"} <= kernels
    per = m["per_env_step"]
    total = sum(v["flop32"] for v in m["kernels"].values()) / (m["envs"] * m["steps"])
    assert abs(total - per["flop32"]) < 1e-6 * per["flop32"]
    assert 1e5 < per["flop32"] < bench.FLOP_CONVENTION_PER_ENV_STEP          # measured work is below the 1 MFLOP convention
    assert 2e4 < per["warp_inst"] < 2e5 and per["flop64"] < 0.05 * per["flop32"]
    light = [v for k, v in m["kernels"].items() if "solve_light" in k][0]
    assert 0.5 * bench.PHASE_ALG_BYTES["solve_light"] * m["envs"] < light["dram_bytes_per_launch"] < 2.0 * bench.PHASE_ALG_BYTES["solve_light"] * m["envs"]
