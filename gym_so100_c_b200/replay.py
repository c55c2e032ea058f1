"""Demonstration replay: runs the action sequences of recorded episodes through the batched env,
one env per episode.

The reference records teleoperated episodes as a pickled list of dicts with the keys ``observations``, ``actions``,
``rewards`` and ``infos`` (scripts/record_teleop.py:177-184, 277-282; merged by scripts/merge_demonstrations.py and
consumed by scripts/train_bc.py:96).  ``replay`` re-simulates all episodes at once on the GPU and returns what the
reference's env would have returned step by step, so a recorded corpus can serve as a regression set (rewards reached,
success) or be re-labelled with another task's reward.

The start state of an episode is the env's reset state with the cube placed
  * where the first recorded observation saw it, if that observation is the 15-float ``so100_state`` vector
    (box position = cube_site = cube centre + 0.01 on every axis for the identity orientation of a reset,
    so100_transfer_cube.xml:13, env.py:137-145), or
  * at ``sample_so100_box_pose(seed + i)`` otherwise (pixel observations carry no cube pose).
"""
from __future__ import annotations

import pickle
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .vec_env import SO100VecEnv, sample_so100_box_pose

CUBE_SITE_OFFSET = 0.01   # so100_transfer_cube.xml:13


def load_demonstrations(path: str) -> List[Dict[str, Any]]:
    """scripts/train_bc.py:96 / scripts/merge_demonstrations.py:13: a pickled list of episode dicts."""
    with open(path, "rb") as f:
        demos = pickle.load(f)
    if isinstance(demos, dict):
        demos = [demos]
    for ep in demos:
        if "actions" not in ep:
            raise ValueError("episode without an 'actions' list")
    return list(demos)


def _first_state_obs(ep: Dict[str, Any]) -> Optional[np.ndarray]:
    obs = ep.get("observations") or []
    if not len(obs):
        return None
    o = obs[0]
    if isinstance(o, dict):
        return None
    o = np.asarray(o, dtype=np.float64).reshape(-1)
    return o if o.shape == (15,) else None


def start_poses(episodes: Sequence[Dict[str, Any]], seed: int = 0) -> np.ndarray:
    """[E, 7] cube poses (xyz + wxyz) the episodes start from."""
    poses = np.zeros((len(episodes), 7), dtype=np.float32)
    for i, ep in enumerate(episodes):
        o = _first_state_obs(ep)
        if o is not None:
            poses[i, :3] = o[:3] - CUBE_SITE_OFFSET
            poses[i, 3] = 1.0
        else:
            poses[i] = sample_so100_box_pose(seed + i)
    return poses


def replay(episodes: Sequence[Dict[str, Any]], task: str = "so100_cube_to_bin", device="cuda:0", seed: int = 0) -> Dict[str, Any]:
    """Re-simulate `episodes` (one env each).  Returns per-step arrays padded to the longest episode:
    ``reward`` [T, E], ``success`` [T, E] (bool), ``obs`` [T, E, 15], ``valid`` [T, E] (step t exists in episode e),
    plus ``episode_return`` [E], ``episode_success`` [E] and, when the recording carries rewards, ``recorded_return`` [E].
    Steps after an episode's end repeat its last action and are masked out by ``valid``."""
    E = len(episodes)
    if E == 0:
        raise ValueError("no episodes")
    lengths = np.array([len(ep["actions"]) for ep in episodes], dtype=np.int64)
    T = int(lengths.max())
    acts = np.zeros((T, E, 6), dtype=np.float32)
    for i, ep in enumerate(episodes):
        a = np.asarray(ep["actions"], dtype=np.float32).reshape(-1, 6)
        acts[:len(a), i] = a
        if len(a) and len(a) < T:
            acts[len(a):, i] = a[-1]
    env = SO100VecEnv(E, task=task, device=device, seed=seed, autoreset=False)
    env.sim.reset(box_pose=torch.from_numpy(start_poses(episodes, seed)))
    reward = np.zeros((T, E), dtype=np.float32)
    success = np.zeros((T, E), dtype=bool)
    obs_out = np.zeros((T, E, 15), dtype=np.float32)
    dev_acts = torch.from_numpy(acts).to(env.device)
    for t in range(T):
        obs, rew, term, trunc, info = env.step(dev_acts[t])
        reward[t] = rew.cpu().numpy()
        success[t] = info["is_success"].cpu().numpy()
        obs_out[t] = obs.cpu().numpy()
    env.close()
    valid = np.arange(T)[:, None] < lengths[None, :]
    out = dict(reward=reward, success=success, obs=obs_out, valid=valid, lengths=lengths,
               episode_return=(reward * valid).sum(axis=0), episode_success=(success & valid).any(axis=0))
    if all("rewards" in ep and len(ep["rewards"]) == len(ep["actions"]) for ep in episodes):
        out["recorded_return"] = np.array([float(np.sum(ep["rewards"])) for ep in episodes])
    return out
