"""Multi-GPU plumbing: one process per GPU, each owning a contiguous env slice; the only
collective on the observation path is ONE all-gather per step of the packed observation
rows onto every rank (the learner reads rank 0's copy).  SURVEY.md 8e.

The reference exchanges no observations (each rank trains on its own envs,
isaacgyminsertion/train.py:58-64; ext_adapt.py:172-178 only all-reduces gradients), so
this is the glue the north star adds, not a replacement of reference code.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Rendezvous from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def env_slice(total_envs, rank, world):
    """Contiguous slice [lo, hi) of global env ids owned by `rank` (remainder to low ranks)."""
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_observations(obs_packed, total_envs=None, out=None):
    """All-gather the packed [n_local, row] observation rows of every rank into
    [total_envs, row] (global env order).  Equal slices use one all_gather_into_tensor;
    ragged slices (total_envs % world != 0) pad to the largest slice."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return obs_packed
    world = dist.get_world_size()
    n_local, row = obs_packed.shape
    if total_envs is None:
        total_envs = n_local * world
    if total_envs % world == 0:
        if out is None:
            out = torch.empty((total_envs, row), dtype=obs_packed.dtype, device=obs_packed.device)
        dist.all_gather_into_tensor(out, obs_packed.contiguous())
        return out
    n_max = -(-total_envs // world)
    padded = torch.zeros((n_max, row), dtype=obs_packed.dtype, device=obs_packed.device)
    padded[:n_local] = obs_packed
    buf = torch.empty((world * n_max, row), dtype=obs_packed.dtype, device=obs_packed.device)
    dist.all_gather_into_tensor(buf, padded)
    parts = []
    for r in range(world):
        lo, hi = env_slice(total_envs, r, world)
        parts.append(buf[r * n_max: r * n_max + (hi - lo)])
    return torch.cat(parts, dim=0)
