"""Host-side job splitter for multi-GPU runs: crops are independent units (SURVEY.md section
8e), so each rank / GPU takes a contiguous block and there is NO collective on the math path.
The only exchange is the final gather of the ``[n_i, max_length] int32`` id rows.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of rank ``rank``: sizes differ by at most one, earlier ranks take the extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_bounds(n_items, world, r)[1] - shard_bounds(n_items, world, r)[0] for r in range(world)]


def gather_ids(local_ids: np.ndarray, n_total: int, group=None, device: Optional[str] = None) -> Optional[np.ndarray]:
    """Final result gather over ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in
    the CPU tests): every rank contributes its ``[n_i, T]`` block, rank 0 returns ``[n_total, T]``
    in the original crop order, other ranks return None.  Blocks are padded to the largest shard
    so one ``all_gather`` of <= 1.2 KB per crop suffices."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.ascontiguousarray(local_ids)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    T = local_ids.shape[1]
    sizes = shard_sizes(n_total, world)
    cap = max(sizes)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    buf = torch.zeros((cap, T), dtype=torch.int32, device=device)
    buf[: local_ids.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_ids, dtype=np.int32)).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    if rank != 0:
        return None
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)], axis=0)
