"""Device-resident SAC+HER rollout loop around :class:`SO100GoalVecEnv`.

Mirrors what the reference's ``scripts/train_sac_her.py:220-254`` builds out of stable_baselines3 parts --
``RecordEpisodeStatistics(SO100GoalEnv())`` inside a ``DummyVecEnv``, feeding ``HerReplayBuffer(n_sampled_goal=4,
goal_selection_strategy="future")`` whose relabelled rewards come from ``env.compute_reward`` -- with everything resident
on the GPU: the replay ring, the episode bookkeeping, the "future" relabelling and the reward recomputation are kernels of
libso100_b200.so (so100_her_begin / so100_her_commit / so100_her_sample), torch only owns the memory.

    env  = SO100GoalVecEnv(65536)
    roll = HerRollout(env, horizon=320, n_sampled_goal=4)
    obs  = roll.reset()
    for _ in range(steps):
        obs, reward, done, info = roll.step(policy(obs))      # stores the transition, auto-resets finished envs
        batch = roll.sample(256)                              # dict of device tensors, 1/5 real + 4/5 relabelled
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import ext


class HerRollout:
    def __init__(self, env, horizon: int = 320, n_sampled_goal: int = 4, seed: int = 0, episode_limit=None):
        """`episode_limit`: the longest episode the caller will produce, if shorter than the env's TimeLimit (tests)."""
        if not getattr(env, "autoreset", False):
            raise ValueError("HerRollout needs an env with autoreset=True (SB3 VecEnv semantics)")
        limit = env.max_episode_steps if episode_limit is None else int(episode_limit)
        if horizon <= limit:
            raise ValueError(f"horizon must exceed the longest episode ({limit} steps): HER samples finished episodes only")
        self.env, self.sim = env, env.sim
        self.T, self.N = int(horizon), env.num_envs
        self.n_sampled_goal = int(n_sampled_goal)
        self.threshold = float(env.distance_threshold)
        self.seed, self.calls, self.pos, self.stored = int(seed), 0, 0, 0
        dev, T, N = env.device, self.T, self.N
        f = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        i = lambda *shape: torch.zeros(shape, dtype=torch.int32, device=dev)
        self.buf = dict(obs=f(T, N, 15), next_obs=f(T, N, 15), achieved=f(T, N, 3), next_achieved=f(T, N, 3), desired=f(T, N, 3),
                        action=f(T, N, 6), reward=f(T, N), done=torch.zeros((T, N), dtype=torch.uint8, device=dev),
                        ep_start=i(T, N), ep_length=i(T, N), cur_start=i(N), cur_length=i(N))
        p = lambda t: C.cast(C.c_void_p(t.data_ptr()), C.c_void_p)
        b = self.buf
        self.ring = ext.HerRing(T, N, *[p(b[k]) for k in ("obs", "next_obs", "achieved", "next_achieved", "desired", "action", "reward",
                                                            "done", "ep_start", "ep_length", "cur_start", "cur_length")])
        self.lib = ext.load()

    def _stream(self):
        return self.sim._stream()

    def reset(self, seed=None):
        obs, _ = self.env.reset(seed=seed)
        for k in ("ep_length", "cur_start", "cur_length"):
            self.buf[k].zero_()
        self.pos = self.stored = 0
        return obs

    def step(self, actions: torch.Tensor):
        """One env step for all envs: the transition (obs, action, reward, next_obs, done) goes into the ring; finished envs
        are reset in the same call (next_obs keeps their terminal observation).  Returns the env's (obs, reward, done, info)."""
        s = self.sim
        a = s._check_in(actions, (self.N, 6))
        vp = lambda t: C.c_void_p(t.data_ptr())
        ext.check(self.lib.so100_her_begin(C.byref(self.ring), self.pos, vp(s.obs), vp(s.achieved), vp(s.desired), vp(a), self._stream()),
                  "so100_her_begin")
        obs, reward, term, trunc, info = self.env.step(a)
        ext.check(self.lib.so100_her_commit(C.byref(self.ring), self.pos, vp(s.obs), vp(s.achieved), vp(s.final_obs), vp(s.reward),
                                            vp(s.terminated), vp(s.truncated), self._stream()), "so100_her_commit")
        self.pos = (self.pos + 1) % self.T
        self.stored = min(self.stored + 1, self.T)
        info["episode"] = {"r": s.ep_return, "l": s.ep_length, "_valid": term | trunc}     # RecordEpisodeStatistics' info["episode"]
        return obs, reward, term | trunc, info

    def sample(self, batch_size: int) -> Dict[str, torch.Tensor]:
        """`batch_size` transitions of finished episodes; the first batch_size // (n_sampled_goal + 1) keep their goal, the rest
        are relabelled ("future").  `index` = (ring position, env, relabelling position or -1); rows of -1 mean "no finished
        episode found" (a buffer that is still filling)."""
        B, dev = int(batch_size), self.env.device
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        out = dict(obs=f(B, 15), action=f(B, 6), next_obs=f(B, 15), achieved=f(B, 3), next_achieved=f(B, 3), desired=f(B, 3), reward=f(B),
                   done=torch.empty(B, dtype=torch.uint8, device=dev), index=torch.empty((B, 3), dtype=torch.int32, device=dev))
        vp = lambda t: C.c_void_p(t.data_ptr())
        self.calls += 1
        ext.check(self.lib.so100_her_sample(C.byref(self.ring), B, self.n_sampled_goal, C.c_float(self.threshold),
                                            C.c_uint64(self.seed & (2**64 - 1)), C.c_uint32(self.calls & 0xFFFFFFFF),
                                            *[vp(out[k]) for k in ("obs", "action", "next_obs", "achieved", "next_achieved", "desired",
                                                                    "reward", "done", "index")], self._stream()), "so100_her_sample")
        return out

    def stats(self) -> Dict[str, float]:
        """Episode statistics since construction (RecordEpisodeStatistics / SB3's ep_rew_mean, ep_len_mean, success_rate)."""
        st = self.sim.episode_stats()
        n = max(st["episodes"], 1)
        return {"episodes": st["episodes"], "ep_rew_mean": st["return_sum"] / n, "ep_len_mean": st["length_sum"] / n,
                "success_rate": st["successes"] / n}
