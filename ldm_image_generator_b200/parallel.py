"""Sharding the sampling batch by image across the GPUs of one box.

Images are independent given (weights, x_T[i], the per-step Python-RNG plan): no BatchNorm/GroupNorm, ChannelNorm is
per pixel (modules.py:24).  So the denoise/decode path needs NO collective -- every rank holds a full weight
replica, seeds Python's ``random`` identically (same expert plan on every rank = the single-device batch's plan),
draws the full noise batch and keeps its contiguous slice.  The only communication is the final gather of the
uint8 images (NCCL all_gather over NVLink; gloo in the CPU tests).
"""
from __future__ import annotations

import random
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the batch owned by `rank`; sizes differ by at most one image."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_noise(shape, seed: int, rank: int, world: int, generator_device: str = "cpu") -> torch.Tensor:
    """x_T of the whole batch drawn from one seeded generator, sliced for this rank -- so the concatenation of all
    shards equals the x_T a single device would have drawn for the same seed."""
    g = torch.Generator(device=generator_device).manual_seed(seed)
    full = torch.randn(*shape, generator=g, device=generator_device)
    lo, hi = shard_bounds(shape[0], rank, world)
    return full[lo:hi].contiguous()


def seed_plan_rng(seed: int) -> None:
    """Every rank must consume the same Python-RNG stream (stochastic depth + expert picks, unet.py:39, modules.py:35)."""
    random.seed(seed)


def gather_images(local: torch.Tensor, global_batch: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank image shards (dim 0) into the full batch, in rank order.  Shards may differ by one image."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(global_batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == mx for lo, hi in sizes):
        out = torch.empty((mx * world,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
