"""Batch collation for ``reader_memory`` instances: what AllenNLP's ``allennlp_collate`` +
``TextField.as_tensor`` + ``move_to_device`` do for this path (SURVEY.md 8b "Tensor input layout").

An instance is a dict ``{"sample1": {"token_ids": [...], "type_ids": [...]}, "label": int|None,
"metadata": {...}}``.  A batch is padded to its own longest sequence:
``sample1 = {"tokens": {"token_ids": i64[B,S], "mask": bool[B,S], "type_ids": i64[B,S]}}``, ``label`` i64[B],
``metadata`` list.  Host tensors are built in pinned memory and copied asynchronously.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch


def collate_instances(instances: List[Dict[str, Any]], device: Optional[torch.device] = None,
                      pad_to: Optional[int] = None, key: str = "sample1", host_only: bool = False) -> Dict[str, Any]:
    """``key``: the text field's name -- ``sample1`` for reader_memory instances, ``sample`` for reader_single ones.
    ``host_only``: build the (pinned) host tensors but leave the copy to the caller (``batch_to_device``): what a
    prefetch thread does."""
    B = len(instances)
    if B == 0:
        raise ValueError("cannot collate an empty batch")
    lens = [len(i[key]["token_ids"]) for i in instances]
    if min(lens) == 0:
        raise ValueError("instance with zero tokens")
    S = max(lens) if pad_to is None else max(pad_to, max(lens))
    pin = device is not None and torch.device(device).type == "cuda" and torch.cuda.is_available()
    ids = torch.zeros(B, S, dtype=torch.int64, pin_memory=pin)
    tids = torch.zeros(B, S, dtype=torch.int64, pin_memory=pin)
    mask = torch.zeros(B, S, dtype=torch.bool, pin_memory=pin)
    for b, inst in enumerate(instances):
        t = inst[key]
        n = lens[b]
        ids[b, :n] = torch.as_tensor(t["token_ids"], dtype=torch.int64)
        if t.get("type_ids") is not None:
            tids[b, :n] = torch.as_tensor(t["type_ids"], dtype=torch.int64)
        mask[b, :n] = True
    batch: Dict[str, Any] = {}
    if device is not None and not host_only:
        ids, tids, mask = (x.to(device, non_blocking=True) for x in (ids, tids, mask))
    batch[key] = {"tokens": {"token_ids": ids, "mask": mask, "type_ids": tids}}
    if all(i.get("label") is not None for i in instances):
        lab = torch.tensor([int(i["label"]) for i in instances], dtype=torch.int64)
        if pin:
            lab = lab.pin_memory()
        batch["label"] = lab.to(device, non_blocking=True) if device is not None and not host_only else lab
    batch["metadata"] = [i["metadata"] for i in instances]
    return batch


def batch_to_device(batch: Dict[str, Any], device: torch.device, key: str = "sample1") -> Dict[str, Any]:
    """Asynchronous H2D of a ``host_only`` batch on the caller's current stream."""
    out = dict(batch)
    out[key] = {"tokens": {k: v.to(device, non_blocking=True) for k, v in batch[key]["tokens"].items()}}
    if "label" in batch:
        out["label"] = batch["label"].to(device, non_blocking=True)
    return out


def batches(instances: List[Dict[str, Any]], batch_size: int):
    """``shuffle: false`` sequential batching (config_memory.json:50-57)."""
    for i in range(0, len(instances), batch_size):
        yield instances[i:i + batch_size]


def plan_length_buckets(lengths: List[int], batch_size: int, window: int = 16) -> List[List[int]]:
    """Batches of instance indices for a mixed-length stream (BASELINE config 5).  The reference pads every batch
    to its longest member in data order, so one 512-token report makes 511 short ones pay for 512
    tokens.  Here each window of ``window * batch_size`` consecutive instances is sorted by length before it is cut
    into batches; results are re-ordered to data order by the caller, so the output is unchanged."""
    out: List[List[int]] = []
    span = max(1, window) * batch_size
    for w0 in range(0, len(lengths), span):
        idx = sorted(range(w0, min(len(lengths), w0 + span)), key=lambda i: (lengths[i], i))
        out.extend(idx[i:i + batch_size] for i in range(0, len(idx), batch_size))
    return out
