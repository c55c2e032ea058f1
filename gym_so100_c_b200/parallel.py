"""Multi-GPU use of the batched env: independent env shards, one process per GPU.

The reference's only parallel axis is "more envs" (SubprocVecEnv workers, scripts/train_sac.py:296-301),
and envs never exchange data, so there is no collective on the step path: rank r of R
owns the global env indices [r*N/R, (r+1)*N/R) and RNG streams are keyed by the GLOBAL index, which
makes trajectories independent of R.  NCCL is used only for the optional episode-statistics all-reduce.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the global env indices owned by `rank` (balanced; first ranks take the remainder)."""
    if not (0 <= rank < world_size) or num_envs_global < 0:
        raise ValueError("bad rank / world_size / env count")
    base, rem = divmod(num_envs_global, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the process group when R > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


STAT_KEYS = ("contact_overflow", "solver_cap_hits", "nonfinite_resets", "episodes", "successes", "newton_iters",
             "solver_runs", "contacts_seen")


def all_reduce_stats(stats: Dict[str, int], device=None) -> Dict[str, int]:
    """Sum the per-shard diagnostics / episode statistics over all ranks (<= 64 bytes, off the step path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(stats.get(k, 0)) for k in STAT_KEYS], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(STAT_KEYS, t.tolist())}


EPISODE_KEYS = ("episodes", "successes", "return_sum", "length_sum")


def all_reduce_episode_stats(stats: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum the episode statistics of BatchedSim.episode_stats() (episodes finished, successes, return sum, length sum: the
    vector SURVEY.md 8e names) over all ranks and add the means a logger reports."""
    vals = [float(stats.get(k, 0.0)) for k in EPISODE_KEYS]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(vals, dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        vals = t.tolist()
    out = dict(zip(EPISODE_KEYS, vals))
    n = max(out["episodes"], 1.0)
    out.update(ep_rew_mean=out["return_sum"] / n, ep_len_mean=out["length_sum"] / n, success_rate=out["successes"] / n)
    return out


def gather_floats(value: float, device=None):
    """[value of rank 0, value of rank 1, ...] on every rank (per-rank timings of a multi-GPU bench line)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(value)]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
