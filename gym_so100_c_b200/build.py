"""Build libso100_b200.so in-tree with nvcc for sm_100a (no torch cpp_extension needed: tensors
cross the C ABI as raw device pointers)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libso100_b200.so")
SOURCES = ["so100_b200.cu"]
HEADERS = ["so100_dev.cuh", "so100_scratch.cuh", "so100_dyn.cuh", "so100_box.cuh", "so100_gjk.cuh", "so100_solve.cuh", "so100_task.cuh", "so100_phases.cuh"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libso100_b200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps += [os.path.join(HERE, "..", "include", f) for f in ("so100_b200.h", "so100_model.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: dict | None = None, out: str | None = None) -> str:
    """`defines` / `out` build tuning variants (e.g. {"SO100_LPE_K3L": 32}) next to the product library;
    point SO100_LIB at one to load it instead."""
    target = out or LIB
    if not force and out is None and not defines and not needs_build():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", target]
    for k, v in (defines or {}).items():
        cmd.append(f"-D{k}={v}")
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return target


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
