"""ctypes binding of libso100_b200.so (the C ABI of include/so100_b200.h).

There is no CPU path: if the CUDA library is missing or no GPU is present, construction fails
loudly (RuntimeError) instead of falling back to anything else.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_lib = None

SYMBOLS = [
    "so100_create", "so100_destroy", "so100_num_envs", "so100_launches_per_step", "so100_reset", "so100_step", "so100_step_host",
    "so100_compute_reward", "so100_get_state", "so100_set_state", "so100_get_aux", "so100_set_aux",
    "so100_substeps", "so100_forward", "so100_diagnostics", "so100_phase_timing", "so100_group_times", "so100_debug_read", "so100_last_error",
    "so100_measure_fp32_peak", "so100_set_episode_outputs", "so100_episode_stats", "so100_graph_stats",
    "so100_her_begin", "so100_her_commit", "so100_her_sample", "so100_render_config", "so100_render",
]

MAX_CONTACTS = 24
NDIAG = 8
TASK_CUBE_TO_BIN = 0
TASK_GOAL = 1
TASK_TOUCH_CUBE = 2
TASK_TOUCH_CUBE_SPARSE = 3


class So100Error(RuntimeError):
    pass


class HerRing(C.Structure):
    """so100_her_ring (include/so100_b200.h): geometry + device pointers of the caller-owned replay ring."""
    _fields_ = [("capacity", C.c_int32), ("num_envs", C.c_int32)] + [(k, C.c_void_p) for k in (
        "obs", "next_obs", "achieved", "next_achieved", "desired", "action", "reward", "done", "ep_start", "ep_length", "cur_start",
        "cur_length")]


def lib_path() -> str:
    return os.environ.get("SO100_LIB") or _build.LIB


def load():
    """Load (never build implicitly on a GPU box: the .so ships in-tree) and type the entry points."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise So100Error(f"{path} is missing: run `python -m gym_so100_c_b200.build` (needs nvcc). "
                         "There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    lib.so100_create.argtypes = [C.c_char_p, C.c_size_t, i32, i32, i32, u64, i64, C.POINTER(vp)]
    lib.so100_destroy.argtypes = [vp]
    lib.so100_num_envs.argtypes = [vp]
    lib.so100_launches_per_step.argtypes = [vp]
    lib.so100_reset.argtypes = [vp] + [vp] * 5 + [vp]
    lib.so100_step.argtypes = [vp, vp, i32] + [vp] * 8 + [vp]
    lib.so100_step_host.argtypes = [vp, vp, i32] + [vp] * 8 + [vp]
    lib.so100_compute_reward.argtypes = [vp, vp, i64, C.c_float, vp, vp]
    lib.so100_get_state.argtypes = [vp] * 6
    lib.so100_set_state.argtypes = [vp] * 6
    lib.so100_get_aux.argtypes = [vp] * 6
    lib.so100_set_aux.argtypes = [vp] * 6
    lib.so100_substeps.argtypes = [vp, i32, vp]
    lib.so100_forward.argtypes = [vp] * 7
    lib.so100_diagnostics.argtypes = [vp, vp, vp]
    lib.so100_phase_timing.argtypes = [vp, i32, vp, vp, vp]
    lib.so100_group_times.argtypes = [vp, vp, vp, vp]
    lib.so100_measure_fp32_peak.argtypes = [i32, vp]
    lib.so100_debug_read.argtypes = [vp, i32, vp, vp, vp]
    lib.so100_set_episode_outputs.argtypes = [vp, vp, vp]
    lib.so100_episode_stats.argtypes = [vp, vp, vp]
    lib.so100_graph_stats.argtypes = [vp, vp, vp, vp]
    lib.so100_render_config.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, i32, C.c_float, C.c_float, i32, i32]
    lib.so100_render.argtypes = [vp, vp, vp]
    ring = C.POINTER(HerRing)
    lib.so100_her_begin.argtypes = [ring, i32, vp, vp, vp, vp, vp]
    lib.so100_her_commit.argtypes = [ring, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.so100_her_sample.argtypes = [ring, i64, i32, C.c_float, u64, C.c_uint32] + [vp] * 9 + [vp]
    lib.so100_last_error.restype = C.c_char_p
    for name in SYMBOLS:
        if name != "so100_last_error":
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().so100_last_error().decode("utf-8", "replace")
        raise So100Error(f"{what} failed ({rc}): {msg}")
