"""Batch sharding over the GPUs of one box (one process per GPU, torch.distributed).

Every roll's chain is independent: no batch statistics, per-clip min-max, per-roll
convolutions.  So the path shards with NO data-path collective; the only exchange is one all-gather of
the finished rolls [B/N,1,T,88] at the end (7.2 MB per rank at B=32/rank), over NCCL on NVLink.
The reference itself has no gather (Lightning DDP ranks just write their own files).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank; sizes differ by at most one (ragged batches allowed)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def all_gather_rolls(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Gather per-rank roll shards [b_r, ...] (contiguous shards from ``shard_bounds``) into [n_total, ...]
    on every rank.  Ragged shards are padded to the largest shard for the collective and trimmed after."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < bmax:
        pad = torch.cat([local, local.new_zeros((bmax - local.shape[0],) + tuple(local.shape[1:]))], 0)
    pad = pad.contiguous()
    out = pad.new_empty((world * bmax,) + tuple(pad.shape[1:]))
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.cat(parts, 0)
    if all(hi - lo == bmax for lo, hi in sizes):
        return out
    return torch.cat([out[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


@torch.no_grad()
def sample_sharded(model, x_T, waveform, noise=None, group=None, generator=None):
    """Run the sampling loop on this rank's contiguous shard of the global batch and all-gather x_0.

    x_T [B,1,T,88], waveform [B,L] and (optional) noise [n,B,1,T,88] are GLOBAL tensors (every rank passes
    the same ones, e.g. drawn from one seeded generator), so an N-rank run returns exactly what a 1-rank
    run returns, sample for sample.  Without pre-drawn noise every rank draws each step's noise for the GLOBAL
    batch (from ``generator``, or its default CUDA generator: seed it alike on every rank) and keeps its slice.
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_bounds(x_T.shape[0], rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    x_loc = x_T[lo:hi].to(dev, non_blocking=True)
    w_loc = waveform[lo:hi].to(dev, non_blocking=True)
    n_loc = None if noise is None else noise[:, lo:hi].to(dev, non_blocking=True)
    x0, spec, _ = model.sample_loop(x_loc, w_loc, noise=n_loc, generator=generator,
                                    shard=None if noise is not None else (x_T.shape[0], lo, hi))
    return all_gather_rolls(x0, x_T.shape[0], group), spec
