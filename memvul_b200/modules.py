"""Parameter containers whose ``state_dict`` keys equal the reference archive's:

    _text_field_embedder.token_embedder_tokens.transformer_model.{embeddings,encoder.layer.N,pooler}.*   (HF BertModel)
    _bert_pooler.pooler.dense.{weight,bias}                      (AllenNLP BertPooler, model_memory.py:64)
    _projector_single._linear_layers.0.{weight,bias}             (AllenNLP FeedForward, model_memory.py:70)
    _projector.weight                                            (nn.Linear(1536, 2, bias=False), model_memory.py:73)

These modules only HOLD fp32 master weights (so ``load_state_dict`` of a reference ``weights.th`` works);
they have no ``forward`` -- all arithmetic runs in the sm_100a kernels via ``memvul_b200.native``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import nn


@dataclass
class BertConfigLite:
    """HF ``BertConfig`` fields the path needs (defaults = bert-base-uncased)."""
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    hidden_act: str = "gelu"
    initializer_range: float = 0.02

    @classmethod
    def from_json_file(cls, path: str) -> "BertConfigLite":
        with open(path, encoding="utf-8") as f:
            raw = json.load(f)
        return cls(**{k: raw[k] for k in cls.__dataclass_fields__ if k in raw})


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("memvul_b200 parameter containers are not callable; the CUDA path computes the forward")


class _Embeddings(_Holder):
    def __init__(self, c: BertConfigLite):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _SelfAttention(_Holder):
    def __init__(self, c):
        super().__init__()
        self.query = nn.Linear(c.hidden_size, c.hidden_size)
        self.key = nn.Linear(c.hidden_size, c.hidden_size)
        self.value = nn.Linear(c.hidden_size, c.hidden_size)


class _DenseLN(_Holder):
    def __init__(self, fan_in, fan_out, eps):
        super().__init__()
        self.dense = nn.Linear(fan_in, fan_out)
        self.LayerNorm = nn.LayerNorm(fan_out, eps=eps)


class _Dense(_Holder):
    def __init__(self, fan_in, fan_out):
        super().__init__()
        self.dense = nn.Linear(fan_in, fan_out)


class _Attention(_Holder):
    def __init__(self, c):
        super().__init__()
        self.self = _SelfAttention(c)
        self.output = _DenseLN(c.hidden_size, c.hidden_size, c.layer_norm_eps)


class _Layer(_Holder):
    def __init__(self, c):
        super().__init__()
        self.attention = _Attention(c)
        self.intermediate = _Dense(c.hidden_size, c.intermediate_size)
        self.output = _DenseLN(c.intermediate_size, c.hidden_size, c.layer_norm_eps)


class _Encoder(_Holder):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(c) for _ in range(c.num_hidden_layers)])


class BertWeights(_Holder):
    """HF ``BertModel`` parameter tree (names only)."""

    def __init__(self, config: BertConfigLite):
        super().__init__()
        self.config = config
        self.embeddings = _Embeddings(config)
        self.encoder = _Encoder(config)
        self.pooler = _Dense(config.hidden_size, config.hidden_size)      # present in archives, unused on this path
        std = config.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                nn.init.normal_(m.weight, mean=0.0, std=std)
                if isinstance(m, nn.Linear):
                    nn.init.zeros_(m.bias)

    @classmethod
    def from_pretrained(cls, path: str) -> "BertWeights":
        """``AutoModel.from_pretrained(path)`` for a local HF BERT directory (config.json +
        pytorch_model.bin), as custom_PTM_embedder.py:99 does, without importing transformers."""
        cfg = BertConfigLite.from_json_file(os.path.join(path, "config.json"))
        model = cls(cfg)
        wfile = os.path.join(path, "pytorch_model.bin")
        sd = torch.load(wfile, map_location="cpu")
        sd = {(k[5:] if k.startswith("bert.") else k): v for k, v in sd.items()}
        sd = {k.replace("LayerNorm.gamma", "LayerNorm.weight").replace("LayerNorm.beta", "LayerNorm.bias"): v
              for k, v in sd.items()}
        own = model.state_dict()
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"{wfile} lacks BERT weights: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        model.load_state_dict({k: sd[k] for k in own})
        return model


class BertPoolerWeights(_Holder):
    """AllenNLP ``BertPooler``: ``self.pooler`` = deep copy of HF BertPooler (dense + tanh)."""

    def __init__(self, hidden: int):
        super().__init__()
        self.pooler = _Dense(hidden, hidden)
        self._embedding_dim = hidden

    def get_output_dim(self) -> int:
        return self._embedding_dim


class FeedForwardWeights(_Holder):
    """AllenNLP ``FeedForward(input_dim, 1, [hidden], ReLU)``: ``_linear_layers.0``."""

    def __init__(self, input_dim: int, hidden: int):
        super().__init__()
        self._linear_layers = nn.ModuleList([nn.Linear(input_dim, hidden)])
        self._output_dim = hidden

    def get_output_dim(self) -> int:
        return self._output_dim


class BasicTextFieldEmbedder(_Holder):
    """AllenNLP ``BasicTextFieldEmbedder``: registers each token embedder as ``token_embedder_<key>``
    and forwards the indexer's tensors to it."""

    def __init__(self, token_embedders: Dict[str, nn.Module]):
        super().__init__()
        self._keys = list(token_embedders.keys())
        for k, emb in token_embedders.items():
            self.add_module(f"token_embedder_{k}", emb)

    def get_output_dim(self) -> int:
        return sum(getattr(self, f"token_embedder_{k}").get_output_dim() for k in self._keys)

    def embedder(self, key: str = "tokens") -> nn.Module:
        return getattr(self, f"token_embedder_{key}")

    def forward(self, text_field_input: Dict[str, Dict[str, torch.Tensor]], **kwargs) -> torch.Tensor:
        if len(self._keys) != 1:
            raise NotImplementedError("memvul_b200 supports the reference's single 'tokens' embedder")
        k = self._keys[0]
        return self.embedder(k)(**text_field_input[k])


def params_version(module: nn.Module) -> int:
    """Changes whenever any parameter is modified in place (load_state_dict, optimiser step, .to())."""
    return hash(tuple((p.data_ptr(), p._version) for p in module.parameters()))
