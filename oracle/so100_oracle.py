"""TEST INFRASTRUCTURE -- ctypes binding of the CPU fp64 oracle (oracle/libso100_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (gym_so100_c_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libso100_oracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "so100_model.h"))
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libso100_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


def build_native() -> str:
    """The CPU-baseline build of the same sources (-O3 -march=native, FMA contraction): always rebuilt on the machine that
    runs it, because a -march=native binary must not travel between hosts.  bench.py's CPU arm loads it with use_native()."""
    subprocess.check_call(["make", "-C", _HERE, "-B", "libso100_oracle_native.so"], stdout=subprocess.DEVNULL)
    return os.path.join(_HERE, "libso100_oracle_native.so")


def use_native():
    """Switch this process to the native build + MuJoCo's solver tolerances (bench.py's CPU arm only; never the checker)."""
    global _lib, _LIB
    _LIB = build_native()
    _lib = None
    lib().so100o_set_solver_mode(1)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.so100o_set_solver_mode.restype = None
        _lib.so100o_set_solver_mode.argtypes = [C.c_int]
        _lib.so100o_create.restype = C.c_void_p
        _lib.so100o_create.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_uint64, C.c_int64]
        _lib.so100o_destroy.argtypes = [C.c_void_p]
        for name in ("so100o_reset", "so100o_substeps", "so100o_forward", "so100o_step", "so100o_set_state",
                     "so100o_get_state", "so100o_set_goal", "so100o_set_counters", "so100o_get_dyn",
                     "so100o_get_contacts", "so100o_get_solver", "so100o_get_efc", "so100o_compute_reward",
                     "so100o_unnormalize"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Batch of independent oracle envs (task 0 = SO100Env cube_to_bin, 1 = SO100GoalEnv)."""

    MAXC = 96

    def __init__(self, blob: bytes, num_envs: int, task: int = 0, seed: int = 0, env_offset: int = 0):
        self._l = lib()
        self.n = num_envs
        self.h = self._l.so100o_create(blob, len(blob), num_envs, task, C.c_uint64(seed), C.c_int64(env_offset))
        if not self.h:
            raise RuntimeError("so100o_create failed (bad model blob)")
        self.total_steps = np.zeros(num_envs, dtype=np.int32)

    def close(self):
        if self.h:
            self._l.so100o_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- env surface
    def reset(self, mask=None, box_pose=None):
        n = self.n
        obs = np.zeros((n, 15), np.float32); ag = np.zeros((n, 3), np.float32); dg = np.zeros((n, 3), np.float32)
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        bp = None if box_pose is None else np.ascontiguousarray(box_pose, dtype=np.float64)
        self._l.so100o_reset(C.c_void_p(self.h), _p(mask), _p(bp), _p(self.total_steps), _p(obs), _p(ag), _p(dg))
        return obs, ag, dg

    def step(self, action, autoreset=False):
        n = self.n
        action = np.ascontiguousarray(action, dtype=np.float32)
        assert action.shape == (n, 6)
        obs = np.zeros((n, 15), np.float32); ag = np.zeros((n, 3), np.float32); dg = np.zeros((n, 3), np.float32)
        rew = np.zeros(n, np.float32); term = np.zeros(n, np.uint8); trunc = np.zeros(n, np.uint8)
        succ = np.zeros(n, np.uint8); fin = np.zeros((n, 15), np.float32)
        self._l.so100o_step(C.c_void_p(self.h), _p(action), int(autoreset), _p(self.total_steps), _p(obs), _p(ag), _p(dg),
                            _p(rew), _p(term), _p(trunc), _p(succ), _p(fin))
        return dict(obs=obs, achieved=ag, desired=dg, reward=rew, terminated=term.astype(bool),
                    truncated=trunc.astype(bool), success=succ.astype(bool), final_obs=fin)

    def substeps(self, nsub=1):
        self._l.so100o_substeps(C.c_void_p(self.h), int(nsub))

    def forward(self):
        self._l.so100o_forward(C.c_void_p(self.h))

    # -- state
    def set_state(self, qpos=None, qvel=None, ctrl=None, warm=None):
        a = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (qpos, qvel, ctrl, warm)]
        self._l.so100o_set_state(C.c_void_p(self.h), *[_p(x) for x in a])

    def get_state(self):
        n = self.n
        qpos = np.zeros((n, 13)); qvel = np.zeros((n, 12)); ctrl = np.zeros((n, 6)); warm = np.zeros((n, 12))
        self._l.so100o_get_state(C.c_void_p(self.h), _p(qpos), _p(qvel), _p(ctrl), _p(warm))
        return qpos, qvel, ctrl, warm

    def set_goal(self, goal):
        g = np.ascontiguousarray(goal, dtype=np.float32)
        self._l.so100o_set_goal(C.c_void_p(self.h), _p(g))

    def set_counters(self, step_count=None, episode=None):
        s = None if step_count is None else np.ascontiguousarray(step_count, dtype=np.int32)
        e = None if episode is None else np.ascontiguousarray(episode, dtype=np.uint32)
        self._l.so100o_set_counters(C.c_void_p(self.h), _p(s), _p(e))

    # -- debug
    def dyn(self, i=0):
        M = np.zeros((12, 12)); bias = np.zeros(12); act = np.zeros(12); qs = np.zeros(12); qacc = np.zeros(12)
        sites = np.zeros((8, 3)); xpos = np.zeros((16, 3)); xquat = np.zeros((16, 4))
        self._l.so100o_get_dyn(C.c_void_p(self.h), i, _p(M), _p(bias), _p(act), _p(qs), _p(qacc), _p(sites), _p(xpos), _p(xquat))
        return dict(M=M, bias=bias, qfrc_act=act, qacc_smooth=qs, qacc=qacc, sites=sites, xpos=xpos, xquat=xquat)

    def contacts(self, i=0):
        geom = np.zeros((self.MAXC, 2), np.int32); data = np.zeros((self.MAXC, 11))
        n = self._l.so100o_get_contacts(C.c_void_p(self.h), i, self.MAXC, _p(geom), _p(data))
        n = min(n, self.MAXC)
        return [dict(geom1=int(geom[c, 0]), geom2=int(geom[c, 1]), dist=data[c, 0], pos=data[c, 1:4].copy(),
                     normal=data[c, 4:7].copy(), force=data[c, 7:11].copy()) for c in range(n)]

    def solver(self, i=0):
        nefc = C.c_int(); it = C.c_int(); g = C.c_double(); ov = C.c_int()
        self._l.so100o_get_solver(C.c_void_p(self.h), i, C.byref(nefc), C.byref(it), C.byref(g), C.byref(ov))
        return dict(nefc=nefc.value, iters=it.value, grad=g.value, overflow=ov.value)

    def efc(self, i=0, maxr=400):
        J = np.zeros((maxr, 12)); aref = np.zeros(maxr); R = np.zeros(maxr); f = np.zeros(maxr); jar = np.zeros(maxr)
        n = self._l.so100o_get_efc(C.c_void_p(self.h), i, maxr, _p(J), _p(aref), _p(R), _p(f), _p(jar))
        return dict(J=J[:n], aref=aref[:n], R=R[:n], force=f[:n], jar=jar[:n])


def test_reward(oracle: Oracle, contacts, cube_site) -> float:
    """single_arm.py:322-380 on a synthetic contact list [(geom1_mjid, geom2_mjid), ...] and cube_site position."""
    l = lib()
    l.so100o_test_reward.restype = C.c_float
    pairs = np.ascontiguousarray(np.asarray(contacts, dtype=np.int32).reshape(-1, 2))
    site = np.ascontiguousarray(cube_site, dtype=np.float64)
    return float(l.so100o_test_reward(C.c_void_p(oracle.h), len(pairs), _p(pairs), _p(site)))


def test_touch_reward(oracle: Oracle, task: int, contacts, cube_site, ee_site) -> float:
    """single_arm.py:149-215 (task 2) / 246-285 (task 3) on a synthetic contact list and cube_site / ee_site positions."""
    l = lib()
    l.so100o_test_touch_reward.restype = C.c_float
    pairs = np.ascontiguousarray(np.asarray(contacts, dtype=np.int32).reshape(-1, 2))
    cs = np.ascontiguousarray(cube_site, dtype=np.float64)
    es = np.ascontiguousarray(ee_site, dtype=np.float64)
    return float(l.so100o_test_touch_reward(C.c_void_p(oracle.h), int(task), len(pairs), _p(pairs), _p(cs), _p(es)))


def test_box_pair(cA, RA, hA, cB, RB, hB):
    """(SAT normal, SAT depth, EPA normal, EPA depth, n_sat_points, epa_hit) for two boxes (row-major R)."""
    l = lib()
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (cA, RA, hA, cB, RB, hB)]
    sat = np.zeros(4); epa = np.zeros(4)
    rc = l.so100o_test_box_pair(*[_p(x) for x in a], _p(sat), _p(epa))
    return sat[:3], sat[3], epa[:3], epa[3], rc // 16, rc % 16


def test_box_manifold(cA, RA, hA, cB, RB, hB):
    """(normal geom1 -> geom2, contact points [n,3], dist [n]) of the oracle's box-box narrow phase (row-major R)."""
    l = lib()
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (cA, RA, hA, cB, RB, hB)]
    nrm = np.zeros(3); pos = np.zeros((8, 3)); dist = np.zeros(8)
    n = l.so100o_test_box_manifold(*[_p(x) for x in a], _p(nrm), _p(pos), _p(dist))
    return nrm, pos[:n].copy(), dist[:n].copy()


test_box_manifold.__test__ = False
test_reward.__test__ = False
test_touch_reward.__test__ = False
test_box_pair.__test__ = False


def set_threads(n: int) -> int:
    """Set the OpenMP team size of the oracle's loops over envs; returns the size in effect."""
    return int(lib().so100o_set_threads(int(n)))


def compute_reward(ag, dg, thr=0.01):
    ag = np.ascontiguousarray(ag, dtype=np.float32).reshape(-1, 3)
    dg = np.ascontiguousarray(dg, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(len(ag), np.float32)
    lib().so100o_compute_reward(_p(ag), _p(dg), len(ag), C.c_float(thr), _p(out))
    return out


def unnormalize(oracle: Oracle, action):
    a = np.ascontiguousarray(action, dtype=np.float32).reshape(-1, 6)
    out = np.zeros_like(a)
    lib().so100o_unnormalize(C.c_void_p(oracle.h), _p(a), len(a), _p(out))
    return out
