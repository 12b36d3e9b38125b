"""``model_single`` (MemVul-m: no external memory) on the sm_100a encoder.

Drop-in for MemVul/model_single.py:36-124 -- same registered name, constructor keywords, output keys
(``meta``, ``probs``, ``loss``) and ``state_dict`` keys (``_projector.0._linear_layers.0.*``,
``_projector.1.weight``).  Forward (:84-92): encoder -> BertPooler -> FeedForward(768->512, ReLU) ->
Linear(512->2, no bias) -> softmax; the head runs as the POOL|HEADER phases of ``memvul_pool_match``
plus ``memvul_single_head``.  The cross-entropy ``loss`` (:93) is evaluated on the host from the
returned logits (inference only; no backward).
"""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np
import torch
from torch import nn

from . import native
from .custom_metric import ClassificationReport
from .model_memory import _tokens, build_text_field_embedder
from .modules import BertPoolerWeights, FeedForwardWeights
from .registrable import Model, Vocabulary


@Model.register("model_single")
class ModelSingle(Model):
    def __init__(self,
                 vocab: Vocabulary,
                 text_field_embedder,
                 PTM: str = "bert-base-uncased",
                 dropout: float = 0.1,
                 label_namespace: str = "class_labels",
                 device: str = "cpu",
                 initializer=None,
                 regularizer=None,
                 *, header_dim: int = 512) -> None:           # 512 in the reference (model_single.py:63); see ModelMemory
        super().__init__(vocab)
        self._device = torch.device(device)
        self._label_namespace = label_namespace
        self._dropout = nn.Dropout(dropout)
        self._idx2token_label = vocab.get_index_to_token_vocabulary(namespace=label_namespace)
        self._idx_pos = vocab.get_token_index(token="pos", namespace=label_namespace)
        self._text_field_embedder = build_text_field_embedder(text_field_embedder)
        self._bert_pooler = BertPoolerWeights(self._text_field_embedder.get_output_dim())
        dim = self._text_field_embedder.get_output_dim()
        self._num_class = self.vocab.get_vocab_size(self._label_namespace)
        if self._num_class != 2:
            raise NotImplementedError("memvul_b200's single head kernel is binary (pos/neg), as in the reference data")
        header = header_dim
        self._projector = nn.Sequential(FeedForwardWeights(dim, header), nn.Linear(header, self._num_class, bias=False))
        self._report = ClassificationReport(self._num_class, self._idx2token_label)
        self._metrics = self._report.parts                        # the reference's attribute name (model_single.py:68)
        if initializer is not None:
            initializer(self)

    def forward(self, sample, label: torch.Tensor = None, metadata: List[Dict[str, Any]] = None) -> Dict[str, Any]:
        output_dict: Dict[str, Any] = dict()
        if metadata:
            output_dict["meta"] = metadata
        t = _tokens(sample)
        if not t["token_ids"].is_cuda:
            raise native.NativeError("memvul_b200 has no CPU path: move the model and the batch to a CUDA device")
        emb = self._text_field_embedder.embedder("tokens")
        hidden = emb(t["token_ids"], t["mask"], t.get("type_ids"), cls_only=True)
        B, S, H = hidden.shape
        fc = self._projector[0]._linear_layers[0]
        head = native.pool_match(hidden, S * H, B, self._bert_pooler.pooler.dense.weight,
                                 self._bert_pooler.pooler.dense.bias, fc.weight, fc.bias,
                                 phase_mask=native.PM_POOL | native.PM_HEADER)
        logits, probs = native.single_head(head["u"], self._projector[1].weight.contiguous())
        emb.check_last_batch()                            # model_single returns host lists, so it syncs anyway (:92)
        probs_h = probs.cpu()
        output_dict["probs"] = probs_h.tolist()
        output_dict["logits_device"] = logits
        if label is not None:
            gold = label.cpu()
            lg = logits.cpu().double()
            lse = torch.logsumexp(lg, dim=-1)
            output_dict["loss"] = (lse - lg[torch.arange(B), gold]).mean().float()
            self._report.update(probs_h.numpy(), gold.numpy())
        return output_dict

    def make_output_human_readable(self, output_dict: Dict[str, Any]):
        """model_single.py:100-110 (note its metadata schema: ``meta["instance"]`` is a dict here)."""
        idx = np.argmax(output_dict["probs"], axis=1)
        rows = []
        for i, k in enumerate(idx):
            inst = output_dict["meta"][i]["instance"]
            inst = inst[0] if isinstance(inst, list) else inst
            rows.append({"Issue_Url": inst["Issue_Url"], "label": inst["label"],
                         "predict": self._idx2token_label[int(k)], "prob": output_dict["probs"][i][self._idx_pos]})
        return rows

    def get_metrics(self, reset: bool = False) -> Dict[str, float]:
        return self._report.report(reset)                         # model_single.py:112-124
