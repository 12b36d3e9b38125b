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


class AsyncGather:
    """The sharded path's collective, off the compute stream.

    Round 1 issued pack (torch.cat) -> all-gather -> unpack on the compute stream after every batch: ~5 tiny launches
    plus NCCL's launch latency, a fixed ~0.4 ms tail per step that was the whole 5 % weak-scaling loss.  Here the shard
    result already is one flat buffer (``ModelMemory.match_batch(..., flat_capacity=max(counts))``,
    ``native.match_flat_layout``), so the exchange is exactly ONE ``all_gather_into_tensor`` with nothing around it, and
    it is enqueued on a side stream that waits for the batch's match kernel: step i's gather overlaps step i+1's encoder.
    ``wait()`` makes the caller's stream wait for the last submitted gather and returns its decoded views."""

    def __init__(self, counts: Sequence[int], device: torch.device, group: Optional[dist.ProcessGroup] = None) -> None:
        self.counts, self.cap, self.group, self.device = list(counts), max(counts), group, device
        self.world = dist.get_world_size(group)
        self.stream = torch.cuda.Stream(device) if torch.device(device).type == "cuda" else None
        self.pending = None

    def submit(self, result: Dict[str, torch.Tensor]) -> None:
        flat = result.get("_flat")
        if flat is None or result.get("_flat_capacity") != self.cap:
            raise ValueError("AsyncGather.submit needs a result produced with flat_capacity=max(counts)")
        G = result["probs"].shape[1]
        out = flat.new_empty(self.world * flat.numel())
        if self.stream is None:                                   # CPU tensors (gloo tests): plain synchronous gather
            dist.all_gather_into_tensor(out, flat, group=self.group)
        else:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                dist.all_gather_into_tensor(out, flat, group=self.group)
            flat.record_stream(self.stream)
            out.record_stream(self.stream)
        self.pending = (out, G, flat.numel())

    def wait(self) -> Optional[Dict[str, List[torch.Tensor]]]:
        if self.pending is None:
            return None
        out, G, n = self.pending
        self.pending = None
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return decode_flat(out, self.counts, self.cap, G, n)


def decode_flat(out: torch.Tensor, counts: Sequence[int], cap: int, G: int, n: int) -> Dict[str, List[torch.Tensor]]:
    """Views (no copies) into the gathered flat buffers: per rank r, probs [counts[r],G,2], best_probs [counts[r],2],
    best_idx [counts[r]] int32."""
    from .native import match_flat_layout
    o_bp, o_bi, n_chk = match_flat_layout(cap, G)
    if n_chk != n:
        raise ValueError("flat buffer size does not match the layout")
    res: Dict[str, List[torch.Tensor]] = {"probs": [], "best_probs": [], "best_idx": []}
    for r, c in enumerate(counts):
        base = out[r * n:(r + 1) * n]
        res["probs"].append(base[:c * G * 2].view(c, G, 2))
        res["best_probs"].append(base[o_bp:o_bp + c * 2].view(c, 2))
        res["best_idx"].append(base[o_bi:o_bi + c].view(torch.int32))
    return res


def gather_match_flat(result: Dict[str, torch.Tensor], counts: Sequence[int],
                      group: Optional[dist.ProcessGroup] = None) -> Dict[str, torch.Tensor]:
    """Synchronous form of AsyncGather for callers that want concatenated tensors: one all-gather of the flat buffer."""
    ag = AsyncGather(counts, result["_flat"].device, group)
    ag.submit(result)
    parts = ag.wait()
    return {k: torch.cat(v) for k, v in parts.items()}


def build_memory_sharded(model, golden_instances: Sequence, chunk: int = 128,
                         group: Optional[dist.ProcessGroup] = None) -> None:
    """the anchor memory of a multi-GPU job is built ONCE -- rank r encodes its contiguous slice of the
    golden instances (in the reference's chunks of 128, predict_memory.py:79-83) and one all-gather of the
    ``[G_r, 512]`` rows gives every rank the whole bank, in file order, with the labels of ALL anchors.  For the
    16,384-anchor stress bank that is 1/world of the 1.58 PFLOP a replicated build costs; rows are bit-identical to a
    single-process build made with the same chunk boundaries inside each slice (every row depends on its own anchor only).
    ``model``: a ``ModelMemory`` (anything with ``forward_on_instances`` filling ``_golden_instances_embeddings`` /
    ``_golden_instances_labels``)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(len(golden_instances), world)
    counts = [e - s for s, e in bounds]
    if min(counts) == 0:
        raise ValueError(f"build_memory_sharded needs at least one golden instance per rank ({len(golden_instances)} over {world})")
    s, e = bounds[rank]
    model._golden_instances_embeddings = None
    model._golden_instances_labels = None
    for c0 in range(s, e, chunk):
        model.forward_on_instances(list(golden_instances[c0:min(e, c0 + chunk)]))
    local = model._golden_instances_embeddings
    local_labels = list(model._golden_instances_labels)
    if local is None or local.shape[0] != counts[rank] or len(local_labels) != counts[rank]:
        raise RuntimeError("build_memory_sharded: the model did not record one bank row per golden instance")
    bank = gather_rows(local.contiguous(), counts, group)
    labels: List[Optional[List[str]]] = [None] * world
    dist.all_gather_object(labels, local_labels, group=group)          # a few KB of CWE ids, once per job
    model._golden_instances_embeddings = bank                          # the setter drops the cached anchor-side match term
    model._golden_instances_labels = [x for part in labels for x in part]
