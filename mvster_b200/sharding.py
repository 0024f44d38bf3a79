"""Multi-GPU partitioning of the forward path.

Two independent axes, both one process per GPU:
  * batch  - frames are independent: replicas, no communication.
  * views  - the epipolar-Transformer aggregation is a plain sum over source views of
             (w_v * cor_v, w_v) followed by one division (mvs4net_utils.py:1054-1060), so a rank
             that owns a subset of the source views produces un-normalised partials and ONE
             all-reduce(sum) per cascade stage over ``[acc | wsum]`` merges them; the division,
             reg2d and the head then run replicated (every rank needs the full-resolution depth
             for the next stage's hypotheses anyway).
``world = replicas x view_parallel``; ranks ``r*P .. r*P+P-1`` form view group ``r``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def partition_views(n_src: int, parts: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced split of source views 1..n_src into ``parts`` shards.
    Returns [(first_view, count)] (1-based view index; count may be 0 when parts > n_src).
    9 views over 4 parts -> 3/2/2/2 (BASELINE cfg4)."""
    base, extra = divmod(n_src, parts)
    out, first = [], 1
    for r in range(parts):
        c = base + (1 if r < extra else 0)
        out.append((first, c))
        first += c
    return out


@dataclass
class ViewShard:
    """This rank's slice of the source views and the process group it reduces over."""
    first_view: int            # 1-based index of the first owned source view
    count: int                 # number of owned source views (may be 0)
    group: Optional[object]    # torch.distributed group of the P ranks sharing one frame (None = default group)
    parts: int = 1

    @property
    def views(self) -> List[int]:
        return list(range(self.first_view, self.first_view + self.count))


def make_view_shard(n_src: int, view_parallel: int, rank: Optional[int] = None, world: Optional[int] = None) -> ViewShard:
    """Build the view group this rank belongs to.  Every rank must call this (new_group is collective)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if world % view_parallel:
        raise ValueError(f"world size {world} is not a multiple of view_parallel={view_parallel}")
    mine = None
    for r0 in range(0, world, view_parallel):
        ranks = list(range(r0, r0 + view_parallel))
        g = dist.new_group(ranks) if view_parallel < world else None
        if rank in ranks:
            mine = g
    first, count = partition_views(n_src, view_parallel)[rank % view_parallel]
    return ViewShard(first, count, mine, view_parallel)


def fused_partials(B: int, D: int, H: int, W: int, G: int, device) -> Tuple[Tensor, Tensor, Tensor]:
    """One flat buffer ``[acc (B*D*H*W*G) | wsum (B*D*H*W)]`` so a stage needs a single collective.
    Returns (flat, acc view [B,D,H,W,G], wsum view [B,D,H,W]); zero-filled (a rank without views
    contributes nothing)."""
    n = B * D * H * W
    flat = torch.zeros(n * (G + 1), dtype=torch.float32, device=device)
    return flat, flat[:n * G].view(B, D, H, W, G), flat[n * G:].view(B, D, H, W)


def sharded_aggregate(partial_fn: Callable[[Tensor, Tensor], None], normalize_fn: Callable[[Tensor, Tensor], Tensor],
                      shape: Tuple[int, int, int, int, int], shard: ViewShard, device) -> Tensor:
    """Run one stage's aggregation view-parallel.
    partial_fn(acc, wsum) must ADD nothing and WRITE this rank's un-normalised partials (it is not
    called when the rank owns no views); normalize_fn(acc, wsum) performs acc / (1e-8 + wsum)."""
    B, D, H, W, G = shape
    flat, acc, wsum = fused_partials(B, D, H, W, G, device)
    if shard.count > 0:
        partial_fn(acc, wsum)
    if shard.parts > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=shard.group)
    return normalize_fn(acc, wsum)


def allreduce_bytes(B: int, D: int, H: int, W: int, G: int) -> int:
    """Payload of the per-stage collective: B*(G+1)*D*H*W fp32."""
    return B * (G + 1) * D * H * W * 4
