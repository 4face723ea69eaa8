"""Host-side sharding logic for multi-GPU runs (one process per GPU, launched by torchrun).

Round 1 shards the path by independent units with no data-path collective:
  * volumes of a scene (each ARaymarchVolume owns its resources) are dealt round-robin to ranks;
  * the rows of one frame can be dealt to ranks in interleaved blocks (the unit tbrm_raymarch_lit renders), which balances
    early ray termination across ranks; gathering the blocks back is plain concatenation.
torch.distributed is used only for barriers, max-over-ranks timing and (optionally) gathering rendered rows.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def volumes_of_rank(n_volumes: int, rank: int, world_size: int) -> List[int]:
    """Volume indices owned by `rank`: round-robin, every volume owned exactly once."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_volumes, world_size))


def row_blocks_of_rank(height: int, rank: int, world_size: int, block_rows: int = 8) -> List[Tuple[int, int]]:
    """Interleaved [begin, end) row blocks of an image of `height` rows owned by `rank`."""
    if block_rows <= 0:
        raise ValueError("block_rows must be positive")
    blocks = [(b, min(b + block_rows, height)) for b in range(0, height, block_rows)]
    return blocks[rank::world_size]


def assemble_rows(height: int, world_size: int, per_rank_rows: Sequence[Sequence], block_rows: int = 8):
    """Inverse of row_blocks_of_rank: per_rank_rows[r] is the list of row-block arrays rank r rendered, in order."""
    import numpy as np

    out = [None] * len(range(0, height, block_rows))
    for r in range(world_size):
        for i, blk in enumerate(per_rank_rows[r]):
            out[r + i * world_size] = blk
    return np.concatenate(out, axis=0)


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction used by bench.py: every multi-GPU number is the max over ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
