"""Seeded synthetic weights and inputs shared by the benchmark, the smoke
test, the parity tests and the CPU oracle.  No pretrained checkpoint, vocabulary or dataset exists in
this environment, so every measurement and parity check runs on these.  Pure data generation: no
model arithmetic lives here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch

EMB = "_text_field_embedder.token_embedder_tokens.transformer_model."


@dataclass(frozen=True)
class BertShape:
    """bert-base-uncased shape (HF BertConfig defaults)."""
    vocab_size: int = 30522
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    intermediate: int = 3072
    max_pos: int = 512
    type_vocab: int = 2
    ln_eps: float = 1e-12
    header: int = 512          # MemVul/model_memory.py:70

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


BERT_BASE = BertShape()
BERT_TINY = BertShape(vocab_size=1024, hidden=128, layers=2, heads=2, intermediate=512,
                      max_pos=512, header=64)
# the same toy encoder under the reference's fixed 512-wide header: what a config-built model (from_params) has
BERT_TINY_H512 = BertShape(vocab_size=1024, hidden=128, layers=2, heads=2, intermediate=512, max_pos=512, header=512)


def synthetic_state_dict(shape: BertShape = BERT_BASE, seed: int = 2021,
                         model: str = "memory") -> Dict[str, torch.Tensor]:
    """Seeded random weights under the reference archive's ``state_dict`` key names
    (SURVEY.md section 8b).  Distribution: HF init N(0, 0.02) for embeddings and
    dense kernels, but Q/K kernels at 0.05 so attention is peaked like a trained
    model's, non-trivial biases / LayerNorm affine so every term is exercised, and
    pooler / header / projector kernels near PyTorch's default Linear scale (0.015-0.02).
    Seed 2021 follows MemVul/config_memory.json:3-8.
    """
    g = torch.Generator().manual_seed(seed)
    H, I = shape.hidden, shape.intermediate

    def n(*size, std=0.02):
        return torch.randn(*size, generator=g, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}
    e = EMB + "embeddings."
    sd[e + "word_embeddings.weight"] = n(shape.vocab_size, H)
    sd[e + "position_embeddings.weight"] = n(shape.max_pos, H)
    sd[e + "token_type_embeddings.weight"] = n(shape.type_vocab, H)
    sd[e + "LayerNorm.weight"] = 1.0 + n(H, std=0.1)
    sd[e + "LayerNorm.bias"] = n(H, std=0.05)
    for l in range(shape.layers):
        p = EMB + f"encoder.layer.{l}."
        sd[p + "attention.self.query.weight"] = n(H, H, std=0.05)
        sd[p + "attention.self.query.bias"] = n(H, std=0.05)
        sd[p + "attention.self.key.weight"] = n(H, H, std=0.05)
        sd[p + "attention.self.key.bias"] = n(H, std=0.05)
        sd[p + "attention.self.value.weight"] = n(H, H)
        sd[p + "attention.self.value.bias"] = n(H)
        sd[p + "attention.output.dense.weight"] = n(H, H)
        sd[p + "attention.output.dense.bias"] = n(H)
        sd[p + "attention.output.LayerNorm.weight"] = 1.0 + n(H, std=0.1)
        sd[p + "attention.output.LayerNorm.bias"] = n(H, std=0.05)
        sd[p + "intermediate.dense.weight"] = n(I, H)
        sd[p + "intermediate.dense.bias"] = n(I)
        sd[p + "output.dense.weight"] = n(H, I)
        sd[p + "output.dense.bias"] = n(H)
        sd[p + "output.LayerNorm.weight"] = 1.0 + n(H, std=0.1)
        sd[p + "output.LayerNorm.bias"] = n(H, std=0.05)
    # HF BertModel's own pooler is present in the archive but unused on this path.
    sd[EMB + "pooler.dense.weight"] = n(H, H)
    sd[EMB + "pooler.dense.bias"] = n(H)
    # AllenNLP BertPooler (model_memory.py:64) and the heads.
    sd["_bert_pooler.pooler.dense.weight"] = n(H, H, std=0.02)
    sd["_bert_pooler.pooler.dense.bias"] = n(H)
    if model == "memory":
        sd["_projector_single._linear_layers.0.weight"] = n(shape.header, H, std=0.02)
        sd["_projector_single._linear_layers.0.bias"] = n(shape.header, std=0.02)
        sd["_projector.weight"] = n(2, 3 * shape.header, std=0.015)      # [Wu | Wv | Wd]
    else:  # model_single.py:62-65
        sd["_projector.0._linear_layers.0.weight"] = n(shape.header, H, std=0.02)
        sd["_projector.0._linear_layers.0.bias"] = n(shape.header, std=0.02)
        sd["_projector.1.weight"] = n(2, shape.header, std=0.03)
    return sd


def synthetic_ids(batch: int, seq: int, lens: Optional[Sequence[int]] = None, seed: int = 2021,
                  vocab_size: int = 30522) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """SURVEY.md section 8d synthetic inputs: ids uniform in [1000, vocab) (or the upper
    half of a tiny vocab), [CLS]=101 first, [SEP]=102 at the last valid slot, pad=0,
    type ids 0, mask = arange(S) < len.  Returns (token_ids i64, mask bool, type_ids i64).
    """
    g = torch.Generator().manual_seed(seed)
    lo = 1000 if vocab_size > 2000 else vocab_size // 2
    ids = torch.randint(lo, vocab_size, (batch, seq), generator=g, dtype=torch.int64)
    if lens is None:
        lens = [seq] * batch
    lens_t = torch.as_tensor(list(lens), dtype=torch.int64)
    assert lens_t.numel() == batch and int(lens_t.min()) >= 2 and int(lens_t.max()) <= seq
    mask = torch.arange(seq)[None, :] < lens_t[:, None]
    ids[:, 0] = 101
    ids[torch.arange(batch), lens_t - 1] = 102
    ids = ids * mask
    return ids, mask, torch.zeros_like(ids)




def load_into(model, sd: Dict[str, torch.Tensor]) -> None:
    """Copy a synthetic (or archive) state dict into a memvul_b200 model; every key must match."""
    own = model.state_dict()
    missing = [k for k in own if k not in sd]
    extra = [k for k in sd if k not in own]
    if missing or extra:
        raise KeyError(f"state_dict mismatch: missing={missing[:4]} unexpected={extra[:4]}")
    model.load_state_dict(sd)


def config_lite(shape: "BertShape"):
    from .modules import BertConfigLite
    return BertConfigLite(vocab_size=shape.vocab_size, hidden_size=shape.hidden, num_hidden_layers=shape.layers,
                          num_attention_heads=shape.heads, intermediate_size=shape.intermediate,
                          max_position_embeddings=shape.max_pos, type_vocab_size=shape.type_vocab,
                          layer_norm_eps=shape.ln_eps)


def build_memory_model(shape: "BertShape" = None, seed: int = 2021, same_first: bool = True, device=None,
                       precision: Optional[str] = None):
    """A ``ModelMemory`` with seeded weights: the object bench.py / smoke() / the parity tests drive.
    ``precision``: the embedder's arithmetic ("fp16" default, "split_fp16" = the opt-in accuracy mode)."""
    from .custom_PTM_embedder import PretrainedTransformerEmbedder
    from .model_memory import ModelMemory
    from .modules import BasicTextFieldEmbedder
    from .registrable import Vocabulary
    shape = shape or BERT_BASE
    vocab = Vocabulary({"labels": ["same", "diff"] if same_first else ["diff", "same"]})
    emb = PretrainedTransformerEmbedder("bert-base-uncased", pretrained_model_path="", config=config_lite(shape),
                                        precision=precision)
    model = ModelMemory(vocab, BasicTextFieldEmbedder({"tokens": emb}), device=str(device or "cpu"), header_dim=shape.header)
    sd = synthetic_state_dict(shape, seed)
    load_into(model, sd)
    model.eval()
    if device is not None:
        model.to(device)
    return model, sd
