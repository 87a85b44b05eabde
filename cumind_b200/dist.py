"""Multi-GPU layout of the search: trees shard by environment, one process per GPU, no collective
in the data path.  The reference has no multi-device code; this is the host-side
bookkeeping around `MCTS.search_batch` / `Agent.select_actions`:

  shard_bounds(total, rank, world)  contiguous slice of the global tree index space owned by `rank`
  gather_rows(local, total, ...)    optional host-side gather of per-tree results (torch.distributed)
Per-tree randomness is keyed by the GLOBAL tree index (`tree_index_base = start`), so the result of a
tree does not depend on how many GPUs the batch is split over.
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of rank's trees: sizes differ by at most one, earlier ranks take the remainder."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} rank={rank} world={world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_bounds(total: int, world: int) -> List[Tuple[int, int]]:
    return [shard_bounds(total, r, world) for r in range(world)]


def gather_rows(local: np.ndarray, total: int, group=None) -> Optional[np.ndarray]:
    """All ranks pass their [n_local, ...] rows; every rank gets the [total, ...] array in global tree
    order.  Works with any initialised torch.distributed backend (NCCL on GPUs, gloo in the CPU tests);
    with no process group it returns `local` unchanged."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = all_bounds(total, world)
    n_max = max(b - a for a, b in bounds)
    a, b = bounds[rank]
    if local.shape[0] != b - a:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {b - a}")
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    pad = np.zeros((n_max,) + local.shape[1:], dtype=local.dtype)
    pad[: b - a] = local
    mine = torch.as_tensor(pad).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.concatenate([p.cpu().numpy()[: bb - aa] for p, (aa, bb) in zip(parts, bounds)], axis=0)
    assert out.shape[0] == total
    return out
