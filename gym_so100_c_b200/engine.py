"""Thin torch-facing wrapper over the C ABI: owns one handle (one GPU), allocates the I/O
tensors once and passes raw device pointers + the current CUDA stream.  torch is plumbing here
(device memory, streams); all arithmetic happens inside libso100_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import ext, model as _model


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def measure_fp32_peak(device: int = 0) -> float:
    """Achieved FP32 FFMA rate of `device` in TFLOP/s (so100_measure_fp32_peak)."""
    out = C.c_float(0.0)
    ext.check(ext.load().so100_measure_fp32_peak(int(device), C.byref(out)), "so100_measure_fp32_peak")
    return float(out.value)


class BatchedSim:
    """N independent bin-a-cube simulations resident on one CUDA device."""

    def __init__(self, num_envs: int, device="cuda:0", task: int = ext.TASK_CUBE_TO_BIN, seed: int = 0,
                 env_offset: int = 0, model_blob: Optional[bytes] = None):
        self.lib = ext.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ext.So100Error("gym_so100_c_b200 runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise ext.So100Error("no CUDA device visible (no CPU fallback)")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self.n = int(num_envs)
        self.task = int(task)
        blob = model_blob if model_blob is not None else _model.pack(_model.load_model())
        self._blob = blob
        h = C.c_void_p()
        ext.check(self.lib.so100_create(blob, len(blob), self.n, index, self.task, C.c_uint64(seed & (2**64 - 1)),
                                        C.c_int64(env_offset), C.byref(h)), "so100_create")
        self.h = h
        n, dev = self.n, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        self.obs = torch.zeros((n, 15), **f32)
        self.final_obs = torch.zeros((n, 15), **f32)
        self.achieved = torch.zeros((n, 3), **f32)
        self.desired = torch.zeros((n, 3), **f32)
        self.reward = torch.zeros(n, **f32)
        self.terminated = torch.zeros(n, **u8)
        self.truncated = torch.zeros(n, **u8)
        self.success = torch.zeros(n, **u8)
        # return / length of the episode an env finished last (RecordEpisodeStatistics' info["episode"]), written by the task kernel
        self.ep_return = torch.zeros(n, **f32)
        self.ep_length = torch.zeros(n, dtype=torch.int32, device=dev)
        ext.check(self.lib.so100_set_episode_outputs(self.h, _ptr(self.ep_return), _ptr(self.ep_length)), "so100_set_episode_outputs")

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "h", None):
            self.lib.so100_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_in(self, t: torch.Tensor, shape, dtype=torch.float32) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t), dtype=dtype)
        t = t.to(device=self.device, dtype=dtype)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t.contiguous()

    # ------------------------------------------------------------------ env surface
    def reset(self, mask: Optional[torch.Tensor] = None, box_pose: Optional[torch.Tensor] = None):
        m = None if mask is None else self._check_in(mask, (self.n,), torch.uint8)
        bp = None if box_pose is None else self._check_in(box_pose, (self.n, 7))
        ext.check(self.lib.so100_reset(self.h, _ptr(m), _ptr(bp), _ptr(self.obs), _ptr(self.achieved), _ptr(self.desired),
                                       self._stream()), "so100_reset")
        return self.obs, self.achieved, self.desired

    def step(self, action: torch.Tensor, autoreset: bool = True, want_final_obs: bool = True):
        a = self._check_in(action, (self.n, 6))
        ext.check(self.lib.so100_step(self.h, _ptr(a), int(autoreset), _ptr(self.obs), _ptr(self.achieved), _ptr(self.desired),
                                      _ptr(self.reward), _ptr(self.terminated), _ptr(self.truncated), _ptr(self.success),
                                      _ptr(self.final_obs) if want_final_obs else None, self._stream()), "so100_step")
        return self.obs, self.reward, self.terminated, self.truncated, self.success

    def step_host(self, action: np.ndarray, autoreset: bool = True) -> Dict[str, np.ndarray]:
        """The C-ABI call a CPU-side user of the reference makes: host buffers in, host buffers out."""
        n = self.n
        a = np.ascontiguousarray(action, dtype=np.float32)
        if a.shape != (n, 6):
            raise ValueError(f"expected shape {(n, 6)}, got {a.shape}")
        if not hasattr(self, "_host"):
            # page-locked result buffers (numpy views of pinned torch tensors): the device -> host copies of so100_step_host
            # then run as true async DMA instead of staged pageable copies
            spec = dict(obs=((n, 15), torch.float32), achieved=((n, 3), torch.float32), desired=((n, 3), torch.float32),
                        reward=((n,), torch.float32), terminated=((n,), torch.uint8), truncated=((n,), torch.uint8),
                        success=((n,), torch.uint8), final_obs=((n, 15), torch.float32))
            self._host_pinned = {k: torch.zeros(shape, dtype=dt).pin_memory() for k, (shape, dt) in spec.items()}
            self._host = {k: v.numpy() for k, v in self._host_pinned.items()}
        o = self._host
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        ext.check(self.lib.so100_step_host(self.h, p(a), int(autoreset), p(o["obs"]), p(o["achieved"]), p(o["desired"]),
                                           p(o["reward"]), p(o["terminated"]), p(o["truncated"]), p(o["success"]),
                                           p(o["final_obs"]), self._stream()), "so100_step_host")
        return o

    def compute_reward(self, achieved: torch.Tensor, desired: torch.Tensor, threshold: float = 0.01) -> torch.Tensor:
        ag = achieved.to(device=self.device, dtype=torch.float32).reshape(-1, 3).contiguous()
        dg = desired.to(device=self.device, dtype=torch.float32).reshape(-1, 3).contiguous()
        if ag.shape != dg.shape:
            raise ValueError("achieved_goal and desired_goal must have the same shape")
        out = torch.empty(ag.shape[0], dtype=torch.float32, device=self.device)
        ext.check(self.lib.so100_compute_reward(_ptr(ag), _ptr(dg), ag.shape[0], C.c_float(threshold), _ptr(out), self._stream()),
                  "so100_compute_reward")
        return out

    # ------------------------------------------------------------------ state
    def get_state(self):
        n, dev = self.n, self.device
        qpos = torch.empty((n, 13), dtype=torch.float32, device=dev)
        qvel = torch.empty((n, 12), dtype=torch.float32, device=dev)
        ctrl = torch.empty((n, 6), dtype=torch.float32, device=dev)
        warm = torch.empty((n, 12), dtype=torch.float32, device=dev)
        ext.check(self.lib.so100_get_state(self.h, _ptr(qpos), _ptr(qvel), _ptr(ctrl), _ptr(warm), self._stream()), "so100_get_state")
        return qpos, qvel, ctrl, warm

    def set_state(self, qpos=None, qvel=None, ctrl=None, warm=None):
        n = self.n
        qpos = None if qpos is None else self._check_in(qpos, (n, 13))
        qvel = None if qvel is None else self._check_in(qvel, (n, 12))
        ctrl = None if ctrl is None else self._check_in(ctrl, (n, 6))
        warm = None if warm is None else self._check_in(warm, (n, 12))
        ext.check(self.lib.so100_set_state(self.h, _ptr(qpos), _ptr(qvel), _ptr(ctrl), _ptr(warm), self._stream()), "so100_set_state")

    def get_aux(self):
        n, dev = self.n, self.device
        goal = torch.empty((n, 3), dtype=torch.float32, device=dev)
        step = torch.empty(n, dtype=torch.int32, device=dev)
        total = torch.empty(n, dtype=torch.int32, device=dev)
        episode = torch.empty(n, dtype=torch.int32, device=dev)
        ext.check(self.lib.so100_get_aux(self.h, _ptr(goal), _ptr(step), _ptr(total), _ptr(episode), self._stream()), "so100_get_aux")
        return goal, step, total, episode

    def set_aux(self, goal=None, step_count=None, total_steps=None, episode=None):
        n = self.n
        goal = None if goal is None else self._check_in(goal, (n, 3))
        step_count = None if step_count is None else self._check_in(step_count, (n,), torch.int32)
        total_steps = None if total_steps is None else self._check_in(total_steps, (n,), torch.int32)
        episode = None if episode is None else self._check_in(episode, (n,), torch.int32)
        ext.check(self.lib.so100_set_aux(self.h, _ptr(goal), _ptr(step_count), _ptr(total_steps), _ptr(episode), self._stream()),
                  "so100_set_aux")

    # ------------------------------------------------------------------ parity / debug
    def substeps(self, nsub: int = 1):
        ext.check(self.lib.so100_substeps(self.h, int(nsub), self._stream()), "so100_substeps")

    def forward(self):
        n, dev = self.n, self.device
        qacc = torch.empty((n, 12), dtype=torch.float32, device=dev)
        ncon = torch.empty(n, dtype=torch.int32, device=dev)
        geom = torch.empty((n, ext.MAX_CONTACTS, 2), dtype=torch.int32, device=dev)
        data = torch.empty((n, ext.MAX_CONTACTS, 11), dtype=torch.float32, device=dev)
        sites = torch.empty((n, 3, 3), dtype=torch.float32, device=dev)
        ext.check(self.lib.so100_forward(self.h, _ptr(qacc), _ptr(ncon), _ptr(geom), _ptr(data), _ptr(sites), self._stream()),
                  "so100_forward")
        return dict(qacc=qacc, ncon=ncon, con_geom=geom, con_data=data, sites=sites)

    def launches_per_step(self) -> int:
        """Kernel launches (graph nodes) one `step` enqueues."""
        return int(self.lib.so100_launches_per_step(self.h))

    def group_times(self):
        """Per-group completion times (ms) of the last step; empty unless SO100_GROUP_TIMES=1 was set at construction."""
        ms = np.zeros(32, dtype=np.float32)
        ng = C.c_int32(0)
        ext.check(self.lib.so100_group_times(self.h, ms.ctypes.data_as(C.c_void_p), C.byref(ng), self._stream()), "so100_group_times")
        return ms[:ng.value].tolist()

    def debug_read(self, what: int) -> torch.Tensor:
        """Raw per-env records (development aid): what=0 state record, what=1 phase workspace."""
        words = C.c_int64(0)
        ext.check(self.lib.so100_debug_read(self.h, int(what), None, C.byref(words), self._stream()), "so100_debug_read")
        out = torch.empty((self.n, int(words.value)), dtype=torch.float32, device=self.device)
        ext.check(self.lib.so100_debug_read(self.h, int(what), _ptr(out), None, self._stream()), "so100_debug_read")
        return out

    def phase_timing(self, enable: bool, read: bool = False):
        """Toggle per-kernel CUDA-event timing of `step`; with read=True returns ({class: ms}, {class: launches})
        accumulated since the previous call (synchronises the stream)."""
        ms = np.zeros(6, dtype=np.float32)
        cnt = np.zeros(6, dtype=np.int32)
        ext.check(self.lib.so100_phase_timing(self.h, int(enable), ms.ctypes.data_as(C.c_void_p) if read else None,
                                              cnt.ctypes.data_as(C.c_void_p) if read else None, self._stream()), "so100_phase_timing")
        names = ["kin_dyn", "collide_box", "solve_light", "task", "collide_hull", "solve_heavy"]
        return dict(zip(names, ms.tolist())), dict(zip(names, cnt.tolist()))

    # ------------------------------------------------------------------ renderer (obs_type "so100_pixels_agent_pos")
    def configure_render(self, width: int, height: int, camera: str = "top"):
        """Scene tables + camera for so100_render (render.py).  Must be called once before `render`."""
        from . import render as _render
        if camera not in _render.CAMERAS:
            raise ValueError(f"camera {camera!r}: one of {sorted(_render.CAMERAS)} (scene_so100.xml:26-29)")
        m = _model.unpack(self._blob)
        planes, adr, num, rgb = _render.scene_tables(m)
        pos, x, y, z = _render.camera_frame(*_render.CAMERAS[camera][:2])
        cam = np.concatenate([pos, x, y, z, [_render.CAMERAS[camera][2]]]).astype(np.float32)
        lights = np.ascontiguousarray(_render.LIGHTS, dtype=np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        ext.check(self.lib.so100_render_config(self.h, p(planes), int(planes.shape[0]), p(adr), p(num), p(rgb), p(cam), p(lights),
                                               int(lights.shape[0]), C.c_float(_render.HEAD_AMBIENT), C.c_float(_render.HEAD_DIFFUSE),
                                               int(width), int(height)), "so100_render_config")
        self.pixels = torch.zeros((self.n, int(height), int(width), 3), dtype=torch.uint8, device=self.device)

    def render(self) -> torch.Tensor:
        """uint8 [N, H, W, 3]: every env's current state seen by the configured camera."""
        if getattr(self, "pixels", None) is None:
            raise ext.So100Error("call configure_render(width, height) first")
        ext.check(self.lib.so100_render(self.h, _ptr(self.pixels), self._stream()), "so100_render")
        return self.pixels

    def episode_stats(self) -> Dict[str, float]:
        """Episodes finished, successes, sum of episode returns and lengths over all envs since construction."""
        out = np.zeros(4, dtype=np.float64)
        ext.check(self.lib.so100_episode_stats(self.h, out.ctypes.data_as(C.c_void_p), self._stream()), "so100_episode_stats")
        return {"episodes": int(out[0]), "successes": int(out[1]), "return_sum": float(out[2]), "length_sum": int(out[3])}

    def graph_stats(self) -> Dict[str, int]:
        c, k, s = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        ext.check(self.lib.so100_graph_stats(self.h, C.byref(c), C.byref(k), C.byref(s)), "so100_graph_stats")
        return {"captures": c.value, "cached": k.value, "staged": s.value}

    def diagnostics(self) -> Dict[str, int]:
        out = np.zeros(ext.NDIAG, dtype=np.int64)
        ext.check(self.lib.so100_diagnostics(self.h, out.ctypes.data_as(C.c_void_p), self._stream()), "so100_diagnostics")
        keys = ["contact_overflow", "solver_cap_hits", "nonfinite_resets", "episodes", "successes", "newton_iters",
                "solver_runs", "contacts_seen"]
        return {k: int(v) for k, v in zip(keys, out)}
