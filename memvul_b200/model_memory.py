"""``model_memory``: MemVul's siamese BERT + external CWE-anchor memory, inference branches, on B200.

Drop-in for MemVul/model_memory.py:39-223: same registered name, constructor keywords, public
methods (``forward``, ``forward_gold_instances``, ``make_output_human_readable``, ``get_metrics``,
``get_output_dim``), externally touched attributes (``_golden_instances_embeddings`` /
``_golden_instances_labels``, reset to ``None`` by MemVul/callbacks.py:48-49) and ``state_dict`` keys.

What runs where
  * ``_instance_forward`` (:90-103)            -> ``memvul_encoder_forward`` + pool/header phases of
                                                  ``memvul_pool_match``
  * ``forward`` test / unlabel branch (:133-147) -> encoder + ONE fused cooperative launch
                                                  (pool -> header -> match -> softmax -> argmax -> gather)
  * ``p.tolist()`` (:143)                       -> one async copy into pinned memory, materialised lazily
  * metric updates (:162-166)                   -> deferred one batch so the host never stalls the stream
The pair-training branch (:149-160) needs a backward pass and is out of scope: it raises.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import native
from .custom_metric import ClassificationReport, SiameseMeasureV1
from .modules import BasicTextFieldEmbedder, BertPoolerWeights, FeedForwardWeights
from .registrable import Model, TokenEmbedder, Vocabulary


class _PinnedPool:
    """Page-locked result buffers are recycled: cudaHostAlloc is a driver-global serialising call (r02c/r02g measured
    end-to-end steps of 8-37 ms instead of 7 when several processes allocated pinned memory every batch)."""

    def __init__(self, keep: int = 8) -> None:
        self._free: Dict[Any, List[torch.Tensor]] = {}
        self._keep = keep

    def get(self, shape, dtype) -> torch.Tensor:
        lst = self._free.get((tuple(shape), dtype))
        return lst.pop() if lst else torch.empty(tuple(shape), dtype=dtype, pin_memory=True)

    def put(self, t: torch.Tensor) -> None:
        lst = self._free.setdefault((tuple(t.shape), t.dtype), [])
        if len(lst) < self._keep:
            lst.append(t)


class _HostBatch:
    """The pinned buffers of one batch; they return to the pool when the last holder (the LazyHostArrays handed to the
    caller, the deferred metric update) lets go."""

    def __init__(self, pool: _PinnedPool, tensors: Dict[str, torch.Tensor]) -> None:
        self.pool, self.tensors, self.leaked = pool, tensors, False

    def __del__(self):
        if not self.leaked:
            for t in self.tensors.values():
                self.pool.put(t)


class LazyHostArray(Sequence):
    """A device result being copied to pinned host memory; behaves like the nested python list the
    reference returns (``len``, iteration, indexing, ``tolist()``) and blocks on first access only."""

    def __init__(self, host: torch.Tensor, event: torch.cuda.Event, check=None, owner: Optional[_HostBatch] = None) -> None:
        self._host, self._event, self._check, self._np, self._owner = host, event, check, None, owner

    def numpy(self) -> np.ndarray:
        if self._np is None:
            self._event.synchronize()
            if self._check is not None:
                self._check()
            view = self._host.numpy()
            if self._owner is None or view.nbytes > (4 << 20):
                if self._owner is not None:
                    self._owner.leaked = True       # a large view may outlive us: its buffer never goes back to the pool
                self._np = view
            else:
                self._np = view.copy()              # small results are copied out so the pinned buffer can be recycled
        return self._np

    def tolist(self) -> list:
        return self.numpy().tolist()

    def __len__(self) -> int:
        return self._host.shape[0]

    def __getitem__(self, i):
        return self.numpy()[i].tolist()

    def __iter__(self):
        return iter(self.tolist())


def build_text_field_embedder(spec) -> nn.Module:
    """Accepts an already-built embedder or the config block
    ``{"token_embedders": {"tokens": {"type": "custom_pretrained_transformer", ...}}}`` (config_memory.json:39-48)."""
    if isinstance(spec, nn.Module):
        return spec
    spec = dict(spec)
    spec.pop("type", None)
    embedders = {k: TokenEmbedder.from_params(dict(v)) for k, v in spec["token_embedders"].items()}
    return BasicTextFieldEmbedder(embedders)


def _tokens(sample: Dict[str, Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    return sample["tokens"] if "tokens" in sample else next(iter(sample.values()))


@Model.register("model_memory")
class ModelMemory(Model):
    def __init__(self,
                 vocab: Vocabulary,
                 text_field_embedder,
                 PTM: str = "bert-base-uncased",
                 dropout: float = 0.1,
                 label_namespace: str = "labels",
                 device: str = "cpu",
                 use_header: bool = True,
                 temperature: float = 1,
                 initializer=None,
                 regularizer=None,
                 *, header_dim: int = 512) -> None:
        # header_dim: the reference hard-codes 512 (model_memory.py:70); keyword-only so that toy-sized test models can
        # shrink it without touching the reference's positional / config signature
        super().__init__(vocab, regularizer)
        self.device = torch.device(device)
        self._use_header = use_header
        self._label_namespace = label_namespace
        self._idx2token_label = self.vocab.get_index_to_token_vocabulary(namespace=label_namespace)
        self._dropout = nn.Dropout(dropout)                      # identity in eval, unused in forward
        self._same_idx = vocab.get_token_index("same", namespace=label_namespace)
        self._text_field_embedder = build_text_field_embedder(text_field_embedder)
        embedding_dim = self._text_field_embedder.get_output_dim()
        self._bert_pooler = BertPoolerWeights(embedding_dim)
        self._num_class = self.vocab.get_vocab_size(self._label_namespace)
        if use_header:
            self._projector_single = FeedForwardWeights(embedding_dim, header_dim)
            embedding_dim = self._projector_single.get_output_dim()
        self._projector = nn.Linear(3 * embedding_dim, 2, bias=False)
        self._temperature = temperature

        self._bank: Optional[torch.Tensor] = None
        self._bank_generation = 0     # bumped by every assignment to _golden_instances_embeddings (incl. reset to None)
        self._golden_instances_labels: Optional[List[str]] = None
        self._vterm = None            # (key, tensor) cache of Wv . bank
        # multi-GPU batch sharding (memvul_b200/dist.py): when set to max(per-rank shard sizes), every result is laid
        # out in one flat buffer that the single all-gather sends as is
        self.shard_capacity: Optional[int] = None

        self._report = ClassificationReport(self._num_class, self._idx2token_label)
        self._metrics = self._report.parts                        # the reference's attribute name (model_memory.py:80)
        self._siamese_metric = SiameseMeasureV1(self._same_idx)
        self._pending: List[Dict[str, Any]] = []      # metric updates waiting for their device->host copy
        self._pinned = _PinnedPool()
        if initializer is not None:
            initializer(self)

    # ------------------------------------------------------------------ helpers
    @property
    def _golden_instances_embeddings(self) -> Optional[torch.Tensor]:
        """The external memory [G, header_dim]; assigned from outside too (MemVul/callbacks.py:48-49 resets it to
        None every epoch).  Every assignment invalidates the cached Wv . bank term: the bank is written through raw
        pointers / torch.cat, so neither its address nor its _version identify its contents."""
        return self._bank

    @_golden_instances_embeddings.setter
    def _golden_instances_embeddings(self, value) -> None:
        self._bank = value
        self._bank_generation += 1
        self._vterm = None

    @property
    def embedder(self):
        return self._text_field_embedder.embedder("tokens")

    def _head_weights(self):
        fc = self._projector_single._linear_layers[0] if self._use_header else None
        return (self._bert_pooler.pooler.dense.weight, self._bert_pooler.pooler.dense.bias,
                None if fc is None else fc.weight, None if fc is None else fc.bias)

    def _encode(self, sample):
        t = _tokens(sample)
        ids, mask = t["token_ids"], t["mask"]
        if not ids.is_cuda:
            raise native.NativeError("memvul_b200 has no CPU path: move the model and the batch to a CUDA device")
        # only hidden[:, 0] is consumed (BertPooler, model_memory.py:99): let the last layer skip the other rows
        hidden = self.embedder(ids, mask, t.get("type_ids"), cls_only=True)
        return hidden, self.embedder.last_bad_mask_flag

    def _instance_forward(self, sample, use_header: bool = False) -> torch.Tensor:
        """model_memory.py:90-103: embed -> BertPooler ([CLS], tanh) -> optional 512-d ReLU header."""
        assert not use_header or hasattr(self, "_projector_single")
        hidden, bad = self._encode(sample)
        B, S, H = hidden.shape
        wp, bp, wh, bh = self._head_weights()
        if use_header:
            out = native.pool_match(hidden, S * H, B, wp, bp, wh, bh, phase_mask=native.PM_POOL | native.PM_HEADER)
            res = out["u"]
        else:
            out = native.pool_match(hidden, S * H, B, wp, bp, None, None, phase_mask=native.PM_POOL)
            res = out["pooled"]
        self._last_bad = bad
        return res

    def forward_gold_instances(self, sample, metadata) -> None:
        """model_memory.py:105-115: append the anchors' feature vectors / labels to the external memory."""
        embedding = self._instance_forward(sample, use_header=self._use_header)
        native.raise_for_flag(int(self._last_bad.item()))
        labels = [m["instance"][0]["label"] for m in metadata]
        if not torch.is_tensor(self._golden_instances_embeddings):
            self._golden_instances_embeddings = embedding
            self._golden_instances_labels = labels
        else:
            self._golden_instances_embeddings = torch.cat([self._golden_instances_embeddings, embedding])
            self._golden_instances_labels.extend(labels)

    def _bank_vterm(self) -> torch.Tensor:
        bank = self._golden_instances_embeddings
        key = (self._bank_generation, tuple(bank.shape), self._projector.weight.data_ptr(),
               self._projector.weight._version)
        if self._vterm is None or self._vterm[0] != key:
            self._vterm = (key, native.bank_prepare(bank.contiguous(), self._projector.weight.contiguous()))
        return self._vterm[1]

    # ------------------------------------------------------------------ forward
    def forward(self,
                sample1=None,
                sample2=None,
                label: torch.Tensor = None,
                metadata: List[Dict[str, Any]] = None) -> Dict[str, Any]:
        output_dict: Dict[str, Any] = dict()
        if metadata and metadata[0]["type"] == "golden":
            self.forward_gold_instances(sample1, metadata)
            return output_dict
        if metadata:
            output_dict["meta"] = metadata
        if not (metadata and metadata[0]["type"] in ["test", "unlabel"]):
            raise NotImplementedError("memvul_b200 implements the inference branches of ModelMemory.forward "
                                      "(metadata type 'golden', 'test', 'unlabel'); pair training is out of scope")
        if not torch.is_tensor(self._golden_instances_embeddings):
            raise RuntimeError("the external memory is empty: run the golden anchors through the model first "
                               "(predict_memory.py:79-83)")
        res = self.match_batch(sample1)
        B = res["best_probs"].shape[0]
        ev = torch.cuda.Event()
        host = {k: self._pinned.get(res[k].shape, res[k].dtype) for k in ("probs", "best_probs", "best_idx")}
        for k in host:
            host[k].copy_(res[k], non_blocking=True)
        flag_h = self._pinned.get((1,), torch.int32)
        flag_h.copy_(res["bad_mask"], non_blocking=True)
        label_h = None
        if label is not None:
            label_h = self._pinned.get(label.shape, label.dtype)
            label_h.copy_(label, non_blocking=True)
        ev.record()
        owner = _HostBatch(self._pinned, dict(host, flag=flag_h, **({"label": label_h} if label_h is not None else {})))

        def check():
            native.raise_for_flag(int(flag_h[0]))

        output_dict["probs"] = LazyHostArray(host["probs"], ev, check, owner)    # list[B][G][2], model_memory.py:143
        output_dict["native"] = {"device": res, "best_probs": LazyHostArray(host["best_probs"], ev, check, owner),
                                 "best_idx": LazyHostArray(host["best_idx"], ev, check, owner)}
        # model_memory.py:162-166 -- metric updates, one batch late so this call never waits for the GPU
        self._flush_metrics()
        self._pending.append({"event": ev, "best_probs": host["best_probs"], "label": label_h,
                              "metadata": metadata, "check": check, "owner": owner})
        return output_dict

    def match_batch(self, sample1, flat_capacity: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Device-side result of the test branch: u, logits/probs [B,G,2], best_idx [B], best_probs [B,2].
        ``flat_capacity``: lay probs / best_probs / best_idx out in one flat buffer (``native.match_flat_layout``) that
        the multi-GPU gather sends without packing."""
        hidden, bad = self._encode(sample1)
        B, S, H = hidden.shape
        wp, bp, wh, bh = self._head_weights()
        if flat_capacity is None:
            flat_capacity = self.shard_capacity
        bank = self._golden_instances_embeddings
        if self._use_header:
            out = native.pool_match(hidden, S * H, B, wp, bp, wh, bh, self._projector.weight, bank.contiguous(),
                                    self._bank_vterm(), same_idx=self._same_idx, phase_mask=native.PM_ALL,
                                    flat_capacity=flat_capacity)
        else:
            # use_header=False (model_memory.py:69,101): the 768-wide pooled vectors are matched directly; the POOL phase
            # writes straight into the match's query buffer and the HEADER phase is skipped
            pooled = torch.empty(B, H, dtype=torch.float32, device=hidden.device)
            out = native.pool_match(hidden, S * H, B, wp, bp, None, None, self._projector.weight, bank.contiguous(),
                                    self._bank_vterm(), same_idx=self._same_idx, u=pooled, pooled=pooled, D=H,
                                    phase_mask=native.PM_POOL | native.PM_UTERM | native.PM_MATCH | native.PM_FINAL,
                                    flat_capacity=flat_capacity)
        out["bad_mask"] = bad
        return out

    def _flush_metrics(self) -> None:
        pend, self._pending = self._pending, []
        for e in pend:
            e["event"].synchronize()
            e["check"]()
            probs = e["best_probs"].numpy()
            if e["label"] is not None:
                gold = e["label"].numpy()
                self._report.update(probs, gold)
            self._siamese_metric(probs, e["metadata"])

    # ------------------------------------------------------------------ outputs
    def make_output_human_readable(self, output_dict: Dict[str, Any]):
        """model_memory.py:169-191: one ``{"Issue_Url", "label", "predict": {cwe_id: P(same)}}`` per sample.
        Anchors sharing a CWE id overwrite each other in bank order (the reference's loop: last wins)."""
        if "meta" not in output_dict or output_dict["meta"][0]["type"] not in ["test", "unlabel"]:
            return output_dict
        labels = self._golden_instances_labels
        last = {}
        for i, name in enumerate(labels):
            last[name] = i
        names = list(set(labels))                      # the reference iterates set(); compare as dicts
        cols = np.asarray([last[n] for n in names], dtype=np.int64)
        probs = output_dict["probs"]
        p = probs.numpy() if isinstance(probs, LazyHostArray) else np.asarray(probs, dtype=np.float32)
        p_same = p[:, cols, self._same_idx].astype(np.float64)     # python floats of fp32 values, as .tolist() gives
        output_dict["predict"] = [dict(zip(names, row)) for row in p_same.tolist()]
        return [{"Issue_Url": meta["instance"][0]["Issue_Url"], "label": meta["instance"][0]["label"],
                 "predict": output_dict["predict"][i]} for i, meta in enumerate(output_dict["meta"])]

    def get_metrics(self, reset: bool = False) -> Dict[str, float]:
        self._flush_metrics()
        metrics = self._report.report(reset)
        if reset:
            s = self._siamese_metric.get_metric(reset)
            metrics["s_precision"], metrics["s_recall"], metrics["s_f1-score"] = s["precision"], s["recall"], s["f1"]
            metrics["s_thres"], metrics["s_auc"] = s["thres"], s["auc"]
            metrics["s_ave_precision_score"] = s["ave_precision_score"]
        return metrics

    def get_output_dim(self, use_header: bool = False) -> int:
        assert not use_header or hasattr(self, "_projector_single")
        if use_header:
            return self._projector_single.get_output_dim()
        return self._text_field_embedder.get_output_dim()
