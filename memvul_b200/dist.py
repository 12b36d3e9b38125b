"""Multi-GPU batch sharding: one process per GPU, weights and anchor bank replicated,
issue reports sharded by rank, ONE all-gather of the per-shard match results per batch.

The reference is single-GPU (predict_memory.py:209-210); this is new functionality.  The path has no
cross-sample term (model_memory.py:133-147), so the only exchange is the result gather: over NCCL
(NVLink 5 / NVSwitch) on GPUs, over gloo in the CPU tests.  Ragged shards (B not divisible by the world
size) are padded to the largest shard for the collective and trimmed afterwards.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [start, end) per rank; the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def balanced_assignment(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-first assignment of samples to ranks by token count (mixed-length streams,
    BASELINE config 5): returns per-rank index lists whose total token cost is near equal.
    Cost model per sample = F(S) of SURVEY.md 8d (linear + quadratic attention term)."""
    def cost(s: int) -> float:
        return 14155776.0 * s + 3072.0 * s * s
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    loads = [0.0] * world
    buckets: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        buckets[r].append(i)
        loads[r] += cost(int(lengths[i]))
    for b in buckets:
        b.sort()
    return buckets


def gather_rows(local: torch.Tensor, counts: Sequence[int], group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather row blocks of possibly different heights: rank r contributes ``local`` [counts[r], ...];
    every rank receives the concatenation [sum(counts), ...].  One collective."""
    world = dist.get_world_size(group)
    if len(counts) != world or local.shape[0] != counts[dist.get_rank(group)]:
        raise ValueError("gather_rows: counts must list every rank's row count")
    mx = max(counts)
    if local.shape[0] < mx:
        pad = local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))
        local = torch.cat([local, pad])
    out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(c == mx for c in counts):
        return out
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)])


def gather_match(result: Dict[str, torch.Tensor], counts: Sequence[int], group: Optional[dist.ProcessGroup] = None,
                 full: bool = True) -> Dict[str, torch.Tensor]:
    """The single collective of the sharded path.  ``full``: all-gather probs [B_r,G,2] (the reference's
    output); otherwise only the reduced form (best_idx, best_probs).  The pieces are packed into one
    fp32 buffer so exactly one all-gather is issued."""
    b = result["best_probs"].shape[0]
    parts = [result["best_idx"].to(torch.float32).view(b, 1), result["best_probs"].view(b, 2)]
    if full:
        parts.append(result["probs"].reshape(b, -1))
    packed = gather_rows(torch.cat(parts, dim=1), counts, group)
    out = {"best_idx": packed[:, 0].round().to(torch.int32), "best_probs": packed[:, 1:3].contiguous()}
    if full:
        g = result["probs"].shape[1]
        out["probs"] = packed[:, 3:].reshape(-1, g, 2)
    return out
