"""Gymnasium-VectorEnv surface of the reference envs, batched on one GPU.

Mirrors (names, argument meaning, returned tuple shapes, error behaviour):
  * ``gym_so100.env.SO100Env(task="so100_cube_to_bin", obs_type="so100_state")``
    (gym_so100/env.py:26-185; registered as gym_so100/SO100CubeToBin-v0 with
    max_episode_steps=700, gym_so100/__init__.py:24-32)            -> :class:`SO100VecEnv`
  * ``gym_so100.env.SO100GoalEnv`` (gym_so100/env.py:188-409)      -> :class:`SO100GoalVecEnv`
  * the SB3 ``VecEnv`` protocol the reference's training scripts consume
    (scripts/train_sac.py:294-301, scripts/train_sac_her.py:220-254) -> :class:`SB3VecEnvAdapter`

All returned arrays are device-resident ``torch`` tensors owned by the env and overwritten by
the next call (the reference returns fresh numpy copies; copy if you keep them).  Declared
deviations: ``obs_type="so100_pixels_agent_pos"`` is drawn by a low-resolution
ray-caster over the collision geometry (render.py), not by MuJoCo's OpenGL renderer, and the
GoalEnv ``"observation"`` entry is the 15-float state vector instead of flattened pixels.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

import numpy as np
import torch

from . import ext
from .engine import BatchedSim
from .model import BOX_RANGE_HI, BOX_RANGE_LO
from .spaces import Box, Dict as DictSpace, batch_box

METADATA = {"render_modes": ["rgb_array"], "render_fps": 50}   # env.py:27


def sample_so100_box_pose(seed=None) -> np.ndarray:
    """gym_so100/utils.py:18-29 restated: MT19937 ``RandomState(seed)``, three uniform draws, identity quat."""
    rng = np.random.RandomState(seed)
    ranges = np.vstack([[BOX_RANGE_LO[0], BOX_RANGE_HI[0]], [BOX_RANGE_LO[1], BOX_RANGE_HI[1]], [BOX_RANGE_LO[2], BOX_RANGE_HI[2]]])
    pos = rng.uniform(ranges[:, 0], ranges[:, 1])
    return np.concatenate([pos, np.array([1, 0, 0, 0])])


class _VecBase:
    metadata = METADATA
    _task = ext.TASK_CUBE_TO_BIN

    def __init__(self, num_envs: int, device="cuda:0", seed: int = 0, env_offset: int = 0,
                 autoreset: bool = True, render_mode: Optional[str] = None):
        if num_envs <= 0:
            raise ValueError("num_envs must be positive")
        self.num_envs = int(num_envs)
        self.render_mode = render_mode
        self.autoreset = bool(autoreset)
        self.sim = BatchedSim(self.num_envs, device=device, task=self._task, seed=seed, env_offset=env_offset)
        self.device = self.sim.device
        self.single_action_space = Box(low=-1, high=1, shape=(6,), dtype=np.float32)        # env.py:75-77
        self.action_space = batch_box(self.single_action_space, self.num_envs)
        self.closed = False

    # -- reference API that has no batched meaning
    def render(self):
        raise NotImplementedError("rendering is out of scope for the batched engine (state observations only)")

    def close(self):
        if not self.closed:
            self.sim.close()
            self.closed = True

    def _box_poses(self, seed) -> Optional[torch.Tensor]:
        """VectorEnv seeding: env i is reset with ``sample_so100_box_pose(seed + i)`` exactly like the
        reference's per-env ``reset(seed=...)``; ``None`` draws on the device (Philox)."""
        if seed is None:
            return None
        seeds = [int(seed) + i for i in range(self.num_envs)] if np.isscalar(seed) else [int(s) for s in seed]
        if len(seeds) != self.num_envs:
            raise ValueError("need one seed per env")
        poses = np.stack([sample_so100_box_pose(s) for s in seeds]).astype(np.float32)
        return torch.from_numpy(poses)

    def _mask(self, options) -> Optional[torch.Tensor]:
        if options and options.get("reset_mask") is not None:
            return torch.as_tensor(options["reset_mask"]).to(torch.uint8)
        return None

    def diagnostics(self) -> Dict[str, int]:
        return self.sim.diagnostics()

    def get_state(self):
        return self.sim.get_state()

    def set_state(self, qpos=None, qvel=None, ctrl=None, warm=None):
        self.sim.set_state(qpos, qvel, ctrl, warm)


class SO100VecEnv(_VecBase):
    """N copies of ``SO100Env(task=..., obs_type="so100_state")`` under the TimeLimit of the registered id
    (gym_so100/__init__.py:4-32): so100_cube_to_bin 700 steps, so100_touch_cube / so100_touch_cube_sparse 300 steps."""

    TASKS = {"so100_cube_to_bin": (ext.TASK_CUBE_TO_BIN, 700), "so100_touch_cube": (ext.TASK_TOUCH_CUBE, 300),
             "so100_touch_cube_sparse": (ext.TASK_TOUCH_CUBE_SPARSE, 300)}

    def __init__(self, num_envs: int, task: str = "so100_cube_to_bin", obs_type: str = "so100_state",
                 observation_width: int = 640, observation_height: int = 480, **kw):
        if task not in self.TASKS:
            raise NotImplementedError(task)           # env.py:117-118
        if obs_type not in ("so100_state", "so100_pixels_agent_pos"):
            raise NotImplementedError(f"obs_type {obs_type!r}")
        self._task, self.max_episode_steps = self.TASKS[task]     # __init__.py:7,17,27
        super().__init__(num_envs, **kw)
        self.task = task
        self.obs_type = obs_type
        self.observation_width, self.observation_height = int(observation_width), int(observation_height)   # env.py:34-35
        if obs_type == "so100_state":
            self.single_observation_space = Box(low=-100.0, high=100.0, shape=(15,), dtype=np.float32)   # env.py:67-73
            self.observation_space = batch_box(self.single_observation_space, self.num_envs)
        else:
            # env.py:50-66: the "top" camera image + the six joint angles.  Drawn by the library's low-resolution ray-caster over
            # the collision geometry (render.py), NOT MuJoCo's OpenGL renderer: use it at the resolutions
            # the reference's own example uses (64 x 48, scripts/example.py:13-14), not to compare pixels with MuJoCo.
            self.sim.configure_render(self.observation_width, self.observation_height, camera="top")
            self.single_observation_space = DictSpace({
                "pixels": Box(low=0, high=255, shape=(self.observation_height, self.observation_width, 3), dtype=np.uint8),
                "agent_pos": Box(low=-10.0, high=10.0, shape=(6,), dtype=np.float32)})
            self.observation_space = DictSpace({k: batch_box(sp, self.num_envs) for k, sp in self.single_observation_space.items()})

    def _format(self, obs):
        """env.py:130-146 on the batch."""
        if self.obs_type == "so100_state":
            return obs
        return {"pixels": self.sim.render(), "agent_pos": obs[:, 9:15]}

    def render(self):
        """Batched ``render()`` (env.py:79-90): uint8 [num_envs, H, W, 3] from the "top" camera at the observation size."""
        if getattr(self.sim, "pixels", None) is None:
            self.sim.configure_render(self.observation_width, self.observation_height, camera="top")
        return self.sim.render()

    def reset(self, seed=None, options: Optional[dict] = None):
        obs, _, _ = self.sim.reset(mask=self._mask(options), box_pose=self._box_poses(seed))
        infos = {"is_success": torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)}   # env.py:169
        return self._format(obs), infos

    def step(self, actions):
        if getattr(actions, "ndim", 2) != 2:
            raise AssertionError("actions must be [num_envs, 6]")   # env.py:173 asserts ndim == 1 per env
        obs, reward, term, trunc, succ = self.sim.step(actions, autoreset=self.autoreset)
        term_b, trunc_b = term.bool(), trunc.bool()
        infos: Dict[str, Any] = {"is_success": succ.bool()}          # env.py:175-177
        if self.autoreset:
            infos["final_obs"] = self.sim.final_obs
            infos["_final_obs"] = term_b | trunc_b
        infos["TimeLimit.truncated"] = trunc_b & ~term_b
        return self._format(obs), reward, term_b, trunc_b, infos


class SO100GoalVecEnv(_VecBase):
    """N copies of ``SO100GoalEnv``: dict observations, sparse 0/-1 reward, 300-step truncation."""

    _task = ext.TASK_GOAL

    def __init__(self, num_envs: int, observation: str = "state", observation_width: int = 640, observation_height: int = 480, **kw):
        """`observation`: "state" (default; the 15-float so100_state vector, the declared deviation that keeps 65536-env
        rollouts feasible) or "pixels" (the reference's own layout, env.py:208-225, 267-270: the flattened "top" image / 255
        followed by the six joint angles, drawn by the library's ray-caster)."""
        if observation not in ("state", "pixels"):
            raise NotImplementedError(f"observation {observation!r}")
        super().__init__(num_envs, **kw)
        self.max_episode_steps = 300                  # env.py:200
        self.distance_threshold = 0.01                # env.py:252
        self.observation_kind = observation
        self.observation_width, self.observation_height = int(observation_width), int(observation_height)
        obs_dim = 15
        if observation == "pixels":
            self.sim.configure_render(self.observation_width, self.observation_height, camera="top")
            obs_dim = self.observation_height * self.observation_width * 3 + 6
        inf = np.inf
        self.single_observation_space = DictSpace({
            "observation": Box(low=-inf, high=inf, shape=(obs_dim,), dtype=np.float32),
            "achieved_goal": Box(low=-inf, high=inf, shape=(3,), dtype=np.float32),   # env.py:228-235
            "desired_goal": Box(low=-inf, high=inf, shape=(3,), dtype=np.float32),
        })
        self.observation_space = DictSpace({k: batch_box(s, self.num_envs) for k, s in self.single_observation_space.items()})

    def _obs(self):
        observation = self.sim.obs
        if self.observation_kind == "pixels":         # env.py:267-270: pixels.flatten() / 255 ++ agent_pos
            pix = self.sim.render().reshape(self.num_envs, -1).to(torch.float32) / 255.0
            observation = torch.cat([pix, self.sim.obs[:, 9:15]], dim=1)
        return {"observation": observation, "achieved_goal": self.sim.achieved, "desired_goal": self.sim.desired}

    def reset(self, seed=None, options: Optional[dict] = None):
        self.sim.reset(mask=self._mask(options), box_pose=self._box_poses(seed))
        infos = {"is_success": torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)}   # env.py:318
        return self._obs(), infos

    def step(self, actions):
        if getattr(actions, "ndim", 2) != 2:
            raise AssertionError("actions must be [num_envs, 6]")
        _, reward, term, trunc, succ = self.sim.step(actions, autoreset=self.autoreset)
        term_b, trunc_b = term.bool(), trunc.bool()
        infos: Dict[str, Any] = {"is_success": succ.bool(), "TimeLimit.truncated": trunc_b}   # env.py:392-403
        if self.autoreset:
            infos["final_obs"] = self.sim.final_obs
            infos["_final_obs"] = term_b | trunc_b
        return self._obs(), reward, term_b, trunc_b, infos

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        """env.py:341-353.  Batched inputs ([..., 3]) -> float32 rewards; single inputs -> Python float."""
        ag = torch.as_tensor(achieved_goal)
        dg = torch.as_tensor(desired_goal)
        r = self.sim.compute_reward(ag, dg, self.distance_threshold)
        if ag.ndim > 1:
            return r.reshape(ag.shape[:-1])
        return float(r[0].item())


class _LazyInfos:
    """The per-env ``infos`` list of an SB3 ``VecEnv.step``: behaves like ``list[dict]`` (len, indexing, iteration) but builds
    env i's dict only when it is asked for, from whole-batch numpy arrays.  Keys as in the reference's stack: ``is_success``,
    ``TimeLimit.truncated`` (env.py:177, 403), and for envs that finished ``terminal_observation`` (SB3 VecEnv auto-reset) and
    ``episode`` = {"r", "l"} (RecordEpisodeStatistics, scripts/train_sac.py:290)."""

    def __init__(self, succ, timeout, done, final, goal_env, prev_desired, ep_return, ep_length):
        self._a = (succ, timeout, done, final, goal_env, prev_desired, ep_return, ep_length)
        self._n = len(done)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        succ, timeout, done, final, goal_env, prev_desired, ep_return, ep_length = self._a
        d = {"is_success": bool(succ[i]), "TimeLimit.truncated": bool(timeout[i])}
        if done[i]:
            if goal_env:
                d["terminal_observation"] = {"observation": final[i].copy(), "achieved_goal": final[i, :3].copy(),
                                             "desired_goal": None if prev_desired is None else prev_desired[i].copy()}
            else:
                d["terminal_observation"] = final[i].copy()
            d["episode"] = {"r": float(ep_return[i]), "l": int(ep_length[i])}
        return d

    def __iter__(self):
        return (self[i] for i in range(self._n))


class SB3VecEnvAdapter:
    """Duck-typed stable_baselines3 ``VecEnv`` over a batched env: numpy in/out, same-step auto-reset,
    ``infos[i]["terminal_observation"]`` and ``env_method("compute_reward", ...)`` for ``HerReplayBuffer``
    (scripts/train_sac_her.py:240-244).  Every call synchronises and copies whole arrays to the host; there is no per-env
    Python on the step path (``infos`` is a lazy sequence).  For rollouts that stay on the device use :class:`her.HerRollout`."""

    def __init__(self, venv: _VecBase):
        if not venv.autoreset:
            raise ValueError("SB3 semantics need autoreset=True")
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space = venv.single_observation_space
        self.action_space = venv.single_action_space
        self._actions = None

    @staticmethod
    def _np(x):
        if isinstance(x, dict):
            return {k: v.cpu().numpy().copy() for k, v in x.items()}
        return x.cpu().numpy().copy()

    def reset(self):
        obs, _ = self.venv.reset()
        out = self._np(obs)
        self._desired = out["desired_goal"] if isinstance(out, dict) else None
        return out

    def seed(self, seed=None):
        self._seed = seed
        return [seed] * self.num_envs

    def step_async(self, actions):
        self._actions = torch.as_tensor(np.asarray(actions, dtype=np.float32))

    def step_wait(self):
        obs, rew, term, trunc, infos = self.venv.step(self._actions)
        done = (term | trunc).cpu().numpy()
        obs_np = self._np(obs)
        sim = self.venv.sim
        # one device -> host copy per array, no per-env Python: the per-env info dicts are built lazily on access
        info_list = _LazyInfos(
            succ=infos["is_success"].cpu().numpy(), timeout=infos["TimeLimit.truncated"].cpu().numpy(), done=done,
            final=infos["final_obs"].cpu().numpy(), goal_env=isinstance(obs, dict),
            prev_desired=getattr(self, "_desired", None),       # the goal of the episode that just ended (a reset draws a new one)
            ep_return=sim.ep_return.cpu().numpy(), ep_length=sim.ep_length.cpu().numpy())
        if isinstance(obs_np, dict):
            self._desired = obs_np["desired_goal"]
        return obs_np, rew.cpu().numpy().copy(), done, info_list

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def env_method(self, method_name: str, *args, indices: Optional[Sequence[int]] = None, **kwargs):
        n = self.num_envs if indices is None else len(list(indices))
        res = getattr(self.venv, method_name)(*args, **kwargs)
        if torch.is_tensor(res):
            res = res.cpu().numpy()
        return [res] * n

    def get_attr(self, name: str, indices=None):
        n = self.num_envs if indices is None else len(list(indices))
        return [getattr(self.venv, name)] * n

    def set_attr(self, name: str, value, indices=None):
        setattr(self.venv, name, value)

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    def close(self):
        self.venv.close()


def make(env_id: str, num_envs: int, obs_type: str = "so100_pixels_agent_pos", **kw):
    """Batched counterpart of ``gym.make`` for the registered ids (gym_so100/__init__.py:4-32).  The registered default
    ``obs_type`` is the pixel observation (__init__.py:11,21,31), which this engine does not render: like the reference's
    kwargs it has to be overridden explicitly, ``make(id, n, obs_type="so100_state")``; the default raises NotImplementedError
    rather than quietly handing out a different observation."""
    ids = {"SO100CubeToBin-v0": "so100_cube_to_bin", "SO100TouchCube-v0": "so100_touch_cube",
           "SO100TouchCubeSparse-v0": "so100_touch_cube_sparse"}
    name = env_id.split("/")[-1]
    if name in ids:
        return SO100VecEnv(num_envs, task=ids[name], obs_type=obs_type, **kw)
    if name in ("SO100Goal-v0", "SO100GoalEnv"):
        return SO100GoalVecEnv(num_envs, **kw)          # the GoalEnv class is not registered and has no obs_type (env.py:191-198)
    raise NotImplementedError(f"{env_id}: not one of the registered ids (gym_so100/__init__.py:4-32) or the GoalEnv")
