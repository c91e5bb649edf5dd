"""Data-parallel plumbing of the hot path (one process per GPU, torch.distributed / NCCL over NVLink).

The path shards over utterances only: every rank holds a replica of the frozen towers, takes a
contiguous slice of the global batch, normalises its CE by the GLOBAL number of label tokens and contributes to
ONE sum-all-reduce of the flat projector gradient per step.  No other collective is on the path.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Contiguous split by sample index; the global batch must divide evenly (config 3: 64 -> 8 x 8)."""
    n = next(iter(batch.values())).shape[0]
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return {k: v[rank * per:(rank + 1) * per] for k, v in batch.items()}


def global_num_items(labels: torch.Tensor, group=None, device: Optional[torch.device] = None) -> int:
    """Number of non-ignored label tokens over all ranks (HF:trainer.py:2133-2143: `num_items_in_batch` gathered and
    summed when average_tokens_across_devices)."""
    n = int((labels[..., 1:] != -100).sum()) if labels.dim() > 1 else int((labels != -100).sum())
    rank, world = world_info(group)
    if world == 1:
        return n
    t = torch.tensor([n], dtype=torch.int64, device=device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def allreduce_flat_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective of the step: in-place SUM over ranks of the flat gradient buffer."""
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad
