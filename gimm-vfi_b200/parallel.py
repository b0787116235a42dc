"""Multi-GPU: frame pairs are independent units, so they shard
across ranks with NO data-path collective; the only communication is ONE all-gather
of the output frames over NCCL/NVLink (gloo on CPU in the tests).  The reference has
no inference-time distribution (its only all_gather is the training metric helper,
src/utils/dist.py:108-116)."""
import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world). One process per GPU, launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, init_method="env://")
    return rank, local, world


def shard_range(n_pairs: int, rank: int, world: int) -> range:
    """Contiguous block of pairs owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_pairs, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def gather_frames(local: torch.Tensor, n_pairs: int = None) -> torch.Tensor:
    """All-gather of per-rank output frames (b_local, 3, H, W) -> (n_pairs, 3, H, W) in pair order.
    Ranks may own different counts (ragged last shard): shards are padded to the max."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [len(shard_range(n_pairs, r, world)) for r in range(world)] if n_pairs is not None else [local.shape[0]] * world
    mx = max(counts)
    buf = local
    if local.shape[0] < mx:
        buf = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, buf.contiguous())
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)
