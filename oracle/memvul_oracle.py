"""CPU fp32 ORACLE for the MemVul batch-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``memvul_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the reported CPU baseline -- never as the thing shipped.

PARITY PINNED AGAINST THE REFERENCE RUN HERE (scope stated exactly): the reference
(panshengyi/MemVul) ships no tests, golden vectors or fixtures for this path (SURVEY.md F2,
section 8c), so there is nothing of its own to check against; instead the reference ITSELF is
executed in the build container -- ``oracle/make_reference_golden.py`` imports the unmodified
first-party files (MemVul/model_memory.py, MemVul/model_single.py, MemVul/custom_PTM_embedder.py,
MemVul/custom_metric.py, MemVul/reader_memory.py, predict_memory.py) from /root/reference over stand-ins for the absent THIRD-PARTY packages
(``oracle/ref_shim.py``: AllenNLP 2.4.0 layers restated from their documented behaviour, ``overrides``;
``transformers.BertModel`` is 5.5 here instead of the pinned 4.1.0), runs them with the seeded
synthetic weights and commits inputs + outputs as ``tests/golden/ref_*.{npz,json}``.
``tests/test_reference_golden.py`` holds this oracle to those outputs (bank, header output, projector
logits <= 2e-5, probabilities <= 2e-6, readable rows, metrics, cal_metrics), and
``tests/test_parity_gpu.py::test_cuda_path_matches_the_reference_run`` holds the CUDA path to them.
Not pinned by the reference: AllenNLP's own BertPooler / FeedForward / BasicTextFieldEmbedder /
metric classes (stand-ins, see ref_shim.py) and the 4.1.0-vs-5.5 transformers difference.
The oracle is additionally checked
  * against an independent implementation of the same third-party arithmetic
    (``transformers.BertModel`` eager attention; tests/test_oracle.py), and
  * against committed golden vectors it produced itself
    (tests/golden/{tiny_ragged,tiny_same1,base_small}.npz, made by oracle/make_golden.py) so drift is detected.

What is restated (all fp32, plain PyTorch on CPU), with the reference lines:
  * ``bert_encoder``           HF transformers==4.1.0 ``BertModel`` as entered from
                               MemVul/custom_PTM_embedder.py:224-228,235 (float mask,
                               additive -10000 key mask, erf-GELU, LayerNorm eps 1e-12).
  * ``embedder_forward``       MemVul/custom_PTM_embedder.py:199-202 (all-zero type ids
                               are dropped), :215-216, :235 (last_hidden_state).
  * ``instance_forward``       MemVul/model_memory.py:90-103 (embed -> BertPooler
                               tanh(W h[:,0] + b) -> FeedForward 768->512 ReLU).
  * ``build_bank``             MemVul/model_memory.py:105-115 + predict_memory.py:79-83
                               (chunks of 128, torch.cat accumulation).
  * ``match``                  MemVul/model_memory.py:133-147 (expand/cat/abs ->
                               Linear(1536->2, no bias) -> softmax -> argmax over anchors
                               of p[:, :, same_idx] -> gather).
  * ``single_forward``         MemVul/model_single.py:84-92 (MemVul-m head).
  * ``vote_labels``            predict_memory.py:168-177 (max over anchors, >= thres).
  * ``human_readable``         MemVul/model_memory.py:169-191 (output schema).
"""
from __future__ import annotations

import math
import os
import sys
from typing import Dict, List, Optional, Tuple

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200.synthetic import (BERT_BASE, BERT_TINY, EMB, BertShape, synthetic_ids,  # noqa: E402,F401
                                   synthetic_state_dict)


# --------------------------------------------------------------------------- BERT
def _ln(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _gelu_erf(x: torch.Tensor) -> torch.Tensor:
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def bert_encoder(sd: Dict[str, torch.Tensor], token_ids: torch.Tensor, mask_float: torch.Tensor,
                 type_ids: Optional[torch.Tensor] = None, shape: BertShape = BERT_BASE,
                 prefix: str = EMB) -> torch.Tensor:
    """``BertModel(input_ids, attention_mask=mask.float()[, token_type_ids]).last_hidden_state``
    with transformers 4.1.0 semantics.  [B,S] -> fp32 [B,S,H]."""
    B, S = token_ids.shape
    H, nH, dh = shape.hidden, shape.heads, shape.head_dim
    e = prefix + "embeddings."
    if type_ids is None:
        type_ids = torch.zeros_like(token_ids)
    x = (sd[e + "word_embeddings.weight"][token_ids]
         + sd[e + "position_embeddings.weight"][:S][None]
         + sd[e + "token_type_embeddings.weight"][type_ids])
    x = _ln(x, sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], shape.ln_eps)
    ext = (1.0 - mask_float.to(torch.float32))[:, None, None, :] * -10000.0
    for l in range(shape.layers):
        p = prefix + f"encoder.layer.{l}."
        lin = torch.nn.functional.linear
        q = lin(x, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"])
        k = lin(x, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"])
        v = lin(x, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"])
        q = q.view(B, S, nH, dh).transpose(1, 2)
        k = k.view(B, S, nH, dh).transpose(1, 2)
        v = v.view(B, S, nH, dh).transpose(1, 2)
        scores = q @ k.transpose(-1, -2) / math.sqrt(dh) + ext
        ctx = (torch.softmax(scores, dim=-1) @ v).transpose(1, 2).reshape(B, S, H)
        a = lin(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
        x = _ln(a + x, sd[p + "attention.output.LayerNorm.weight"],
                sd[p + "attention.output.LayerNorm.bias"], shape.ln_eps)
        h = _gelu_erf(lin(x, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        o = lin(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
        x = _ln(o + x, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], shape.ln_eps)
    return x


def embedder_forward(sd, token_ids, mask, type_ids=None, shape: BertShape = BERT_BASE) -> torch.Tensor:
    """custom_PTM_embedder.py:199-202 (type ids dropped when all zero; too-large ids raise),
    :215-216/:224-228 (bool mask cast to float), :235."""
    if type_ids is not None:
        mx = int(type_ids.max())
        if mx == 0:
            type_ids = None
        elif mx >= shape.type_vocab:
            raise ValueError("Found type ids too large for the chosen transformer model.")
    return bert_encoder(sd, token_ids, mask.float(), type_ids, shape)


def instance_forward(sd, token_ids, mask, type_ids=None, shape: BertShape = BERT_BASE,
                     use_header: bool = True, model: str = "memory") -> torch.Tensor:
    """model_memory.py:90-103.  Returns u [B,512] (or pooled [B,768] without the header)."""
    hidden = embedder_forward(sd, token_ids, mask, type_ids, shape)
    lin = torch.nn.functional.linear
    pooled = torch.tanh(lin(hidden[:, 0], sd["_bert_pooler.pooler.dense.weight"],
                            sd["_bert_pooler.pooler.dense.bias"]))
    if not use_header:
        return pooled
    key = "_projector_single._linear_layers.0." if model == "memory" else "_projector.0._linear_layers.0."
    return torch.relu(lin(pooled, sd[key + "weight"], sd[key + "bias"]))


def build_bank(sd, anchors: List[Tuple[torch.Tensor, torch.Tensor]], shape: BertShape = BERT_BASE,
               chunk: int = 128) -> torch.Tensor:
    """predict_memory.py:79-83 + model_memory.py:105-115.  ``anchors`` is a list of
    (token_ids [S_g], mask [S_g]) 1-D tensors; each chunk is padded to its own longest
    member (AllenNLP collate) and the bank is the torch.cat of the chunk outputs."""
    bank = None
    for c0 in range(0, len(anchors), chunk):
        part = anchors[c0:c0 + chunk]
        S = max(int(t.numel()) for t, _ in part)
        ids = torch.zeros(len(part), S, dtype=torch.int64)
        msk = torch.zeros(len(part), S, dtype=torch.bool)
        for i, (t, m) in enumerate(part):
            ids[i, :t.numel()] = t
            msk[i, :m.numel()] = m
        emb = instance_forward(sd, ids, msk, torch.zeros_like(ids), shape)
        bank = emb if bank is None else torch.cat([bank, emb])
    return bank


def match(u: torch.Tensor, bank: torch.Tensor, w_proj: torch.Tensor, same_idx: int):
    """model_memory.py:133-147, literally (expand / cat / abs / Linear / softmax / argmax).
    Returns dict(logits [B,G,2], p [B,G,2], best_idx [B] i64, probs [B,2])."""
    B, D = u.shape
    G = bank.shape[0]
    su = u.view(B, -1, D).expand(-1, G, -1)
    gv = bank.expand(B, -1, -1)
    logits = torch.nn.functional.linear(torch.cat([su, gv, torch.abs(su - gv)], -1), w_proj)
    p = torch.softmax(logits, dim=-1)
    idx = torch.argmax(p, dim=1)[:, same_idx]
    probs = torch.stack([p[i][idx[i]] for i in range(B)])
    return {"logits": logits, "p": p, "best_idx": idx, "probs": probs}


def match_separable(u, bank, w_proj, same_idx):
    """SURVEY.md F4 identity: logits = Wu.u + Wv.v + Wd.|u - v|.  The form the CUDA kernel
    uses; kept here so a CPU test pins the identity against ``match``."""
    D = u.shape[1]
    wu, wv, wd = w_proj[:, :D], w_proj[:, D:2 * D], w_proj[:, 2 * D:]
    logits = (u @ wu.T)[:, None, :] + (bank @ wv.T)[None, :, :] \
        + torch.einsum("bgk,ck->bgc", (u[:, None, :] - bank[None, :, :]).abs(), wd)
    p = torch.softmax(logits, dim=-1)
    idx = torch.argmax(p[:, :, same_idx], dim=1)
    return {"logits": logits, "p": p, "best_idx": idx, "probs": p[torch.arange(u.shape[0]), idx]}


def memory_forward(sd, token_ids, mask, type_ids, bank, same_idx: int,
                   shape: BertShape = BERT_BASE):
    """ModelMemory.forward test/unlabel branch end to end (model_memory.py:133-147)."""
    u = instance_forward(sd, token_ids, mask, type_ids, shape)
    out = match(u, bank, sd["_projector.weight"], same_idx)
    out["u"] = u
    return out


def single_forward(sd, token_ids, mask, type_ids=None, shape: BertShape = BERT_BASE):
    """ModelSingle.forward (model_single.py:84-92): logits [B,2], probs [B,2]."""
    h = instance_forward(sd, token_ids, mask, type_ids, shape, model="single")
    logits = torch.nn.functional.linear(h, sd["_projector.1.weight"])
    return {"logits": logits, "probs": torch.softmax(logits, -1)}


# --------------------------------------------------------------------------- host rules
def vote_labels(p_same: torch.Tensor, thres: float = 0.5):
    """predict_memory.py:170-177: vote_prob = max over anchors; 'pos' iff >= thres."""
    vote = p_same.max(dim=1).values
    return vote, ["pos" if float(v) >= thres else "neg" for v in vote]


def human_readable(p: torch.Tensor, bank_labels: List[str], metadata: List[dict], same_idx: int):
    """model_memory.py:169-191: one {"Issue_Url","label","predict":{cwe: p_same}} per sample.
    Anchors sharing a CWE id overwrite each other in bank order (last wins), as in the
    reference's ``vote_num[golden_name] = p[idx_same]`` loop."""
    rows = []
    pl = p.tolist()
    for probs, meta in zip(pl, metadata):
        vote = {c: 0 for c in set(bank_labels)}
        for pr, name in zip(probs, bank_labels):
            vote[name] = pr[same_idx]
        rows.append({"Issue_Url": meta["instance"][0]["Issue_Url"],
                     "label": meta["instance"][0]["label"], "predict": vote})
    return rows


def flops_per_issue(seq_len: int) -> float:
    """F(S) = 12*(14,155,776*S + 3,072*S^2) + 2*768^2 + 2*768*512."""
    return 12.0 * (14155776.0 * seq_len + 3072.0 * seq_len * seq_len) + 2 * 768 ** 2 + 2 * 768 * 512
