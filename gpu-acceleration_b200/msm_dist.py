"""Multi-GPU plumbing for the MSM: point-range sharding and the one tiny collective.

MSM shards naturally: rank g owns the contiguous point range
[g*ceil(n/G), min(n, (g+1)*ceil(n/G))) of bases AND scalars, runs the complete single-GPU
pipeline on it, and produces one 96-byte Jacobian partial.  The only exchange step is an
all-gather of those partials (96 B x world, `torch.distributed`, NCCL on GPUs / gloo in the CPU
tests); every rank then adds the G partials (device kernel `b200msm_sum_partials_device` in the
product; an injected adder in the CPU tests).  Bases never move between GPUs.

The reference has no multi-device path at all (single `Device::system_default()`,
/root/reference/mopro-msm/src/msm/metal_msm/host/gpu.rs:3-5); this attaches where its CPU
`final_reduction` (metal_msm.rs:204-262) ends.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Same split as the C library's shard_ranges(): ceil-div contiguous ranges; trailing ranks may be empty."""
    per = (n + world - 1) // world
    b = min(n, rank * per)
    e = min(n, (rank + 1) * per)
    return b, e


def all_gather_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """partial: (96,) uint8 tensor (device for NCCL, CPU for gloo) -> (world*96,) uint8, rank-major."""
    world = dist.get_world_size(group)
    assert partial.dtype == torch.uint8 and partial.numel() == 96
    out = torch.empty(world * 96, dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    return out


def combine(partial: torch.Tensor, sum_fn: Callable[[torch.Tensor, int], torch.Tensor], group=None) -> torch.Tensor:
    """All-gather the per-rank partial sums and add them with `sum_fn(gathered, world) -> (96,) uint8`."""
    world = dist.get_world_size(group)
    gathered = all_gather_partials(partial, group)
    return sum_fn(gathered, world)
