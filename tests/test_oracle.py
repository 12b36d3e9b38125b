"""CPU tests of the oracle (the checker itself).  The reference has no tests or golden vectors
(SURVEY.md F2), so the oracle is pinned by (1) an independent implementation of the third-party
arithmetic -- HF ``transformers.BertModel`` with eager attention -- and (2) committed golden vectors."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import memvul_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hf_model(sd, shape):
    transformers = pytest.importorskip("transformers")
    cfg = transformers.BertConfig(vocab_size=shape.vocab_size, hidden_size=shape.hidden, num_hidden_layers=shape.layers,
                                  num_attention_heads=shape.heads, intermediate_size=shape.intermediate,
                                  max_position_embeddings=shape.max_pos, type_vocab_size=shape.type_vocab,
                                  layer_norm_eps=shape.ln_eps, hidden_act="gelu", attn_implementation="eager")
    m = transformers.BertModel(cfg, add_pooling_layer=True).eval()
    own = m.state_dict()
    src = {k[len(O.EMB):]: v for k, v in sd.items() if k.startswith(O.EMB)}
    missing = [k for k in own if k not in src and not k.endswith("position_ids")]
    assert not missing, missing
    m.load_state_dict({k: src[k] for k in own if k in src}, strict=False)
    return m


@pytest.mark.parametrize("shape,B,S,lens", [(O.BERT_TINY, 4, 40, [40, 7, 22, 2]), (O.BERT_BASE, 2, 24, [24, 10])])
def test_oracle_bert_matches_hf_bertmodel(shape, B, S, lens):
    """custom_PTM_embedder.py:224-235: BertModel(input_ids, attention_mask=mask.float()).last_hidden_state."""
    sd = O.synthetic_state_dict(shape)
    ids, mask, tids = O.synthetic_ids(B, S, lens=lens, vocab_size=shape.vocab_size)
    with torch.no_grad():
        ours = O.embedder_forward(sd, ids, mask, tids, shape)
        hf = _hf_model(sd, shape)(input_ids=ids, attention_mask=mask.float()).last_hidden_state
    # padded query rows differ between -10000 (4.1.0) and dtype-min (5.x) masking only where every key is masked: never
    assert float((ours - hf)[mask].abs().max()) < 2e-5
    assert float((ours[:, 0] - hf[:, 0]).abs().max()) < 2e-5


def test_type_ids_semantics():
    """custom_PTM_embedder.py:199-206: all-zero type ids are dropped; non-zero are used; too large raises."""
    shape = O.BERT_TINY
    sd = O.synthetic_state_dict(shape)
    ids, mask, tids = O.synthetic_ids(2, 16, vocab_size=shape.vocab_size)
    a = O.embedder_forward(sd, ids, mask, tids, shape)
    b = O.embedder_forward(sd, ids, mask, None, shape)
    assert torch.equal(a, b)
    t1 = tids.clone(); t1[:, 8:] = 1
    assert not torch.allclose(O.embedder_forward(sd, ids, mask, t1, shape), a)
    t2 = tids.clone(); t2[0, 0] = 2
    with pytest.raises(ValueError):
        O.embedder_forward(sd, ids, mask, t2, shape)


def test_separable_identity_and_argmax_rule():
    """concat form == Wu.u + Wv.v + Wd.|u-v|; first maximum wins; same_idx selects the column."""
    g = torch.Generator().manual_seed(3)
    u = torch.relu(torch.randn(7, 512, generator=g))
    bank = torch.relu(torch.randn(129, 512, generator=g))
    w = torch.randn(2, 1536, generator=g) * 0.03
    for same in (0, 1):
        a, b = O.match(u, bank, w, same), O.match_separable(u, bank, w, same)
        assert float((a["logits"] - b["logits"]).abs().max()) < 2e-6
        assert torch.equal(a["best_idx"], b["best_idx"])
        assert torch.equal(a["best_idx"], a["p"][:, :, same].argmax(1))
        assert torch.allclose(a["probs"], a["p"][torch.arange(7), a["best_idx"]])
    # exact ties (duplicated anchors): lowest index
    bank2 = bank[:1].repeat(6, 1)
    assert O.match(u, bank2, w, 0)["best_idx"].tolist() == [0] * 7
    # G = 1
    one = O.match(u, bank[:1], w, 0)
    assert one["logits"].shape == (7, 1, 2) and one["best_idx"].tolist() == [0] * 7


def test_bank_chunking_equals_single_pass():
    """predict_memory.py:81-83: 128 anchors then the rest == all at once (row-wise independent)."""
    shape = O.BERT_TINY
    sd = O.synthetic_state_dict(shape)
    n = 131
    lens = [int(x) for x in torch.randint(2, 24, (n,), generator=torch.Generator().manual_seed(1))]
    a_ids, a_mask, _ = O.synthetic_ids(n, 24, lens=lens, vocab_size=shape.vocab_size)
    anchors = [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(n)]
    with torch.no_grad():
        chunked = O.build_bank(sd, anchors, shape, chunk=128)
        once = O.build_bank(sd, anchors, shape, chunk=1000)
    assert chunked.shape == (n, shape.header)
    assert float((chunked - once).abs().max()) < 1e-5      # padding length differs between the two -> fp noise only


def test_padding_invariance():
    """Batches are padded to their longest member; extra padding must not change valid rows."""
    shape = O.BERT_TINY
    sd = O.synthetic_state_dict(shape)
    ids, mask, tids = O.synthetic_ids(2, 20, lens=[20, 11], vocab_size=shape.vocab_size)
    ids2 = torch.cat([ids, torch.zeros(2, 13, dtype=torch.int64)], 1)
    mask2 = torch.cat([mask, torch.zeros(2, 13, dtype=torch.bool)], 1)
    with torch.no_grad():
        a = O.instance_forward(sd, ids, mask, tids, shape)
        b = O.instance_forward(sd, ids2, mask2, torch.zeros_like(ids2), shape)
    assert float((a - b).abs().max()) < 1e-5


def test_vote_and_readable_schema():
    """predict_memory.py:168-177 and model_memory.py:169-191."""
    p = torch.tensor([[[0.2, 0.8], [0.7, 0.3], [0.6, 0.4]], [[0.49, 0.51], [0.1, 0.9], [0.5, 0.5]]])
    vote, lab = O.vote_labels(p[:, :, 0], 0.5)
    assert vote.tolist() == pytest.approx([0.7, 0.5]) and lab == ["pos", "pos"]
    assert O.vote_labels(p[:, :, 0], 0.6)[1] == ["pos", "neg"]
    meta = [{"type": "unlabel", "instance": [{"label": "neg", "Issue_Url": "a"}]},
            {"type": "unlabel", "instance": [{"label": "CWE-79", "Issue_Url": "b"}]}]
    rows = O.human_readable(p, ["CWE-1", "CWE-2", "CWE-1"], meta, 0)
    assert rows[0] == {"Issue_Url": "a", "label": "neg", "predict": {"CWE-1": pytest.approx(0.6), "CWE-2": pytest.approx(0.7)}}
    assert rows[1]["label"] == "CWE-79" and set(rows[1]["predict"]) == {"CWE-1", "CWE-2"}


def test_flops_formula():
    assert O.flops_per_issue(512) == pytest.approx(96.64e9, rel=1e-3)
    assert O.flops_per_issue(128) == pytest.approx(22.35e9, rel=1e-3)


@pytest.mark.parametrize("name", ["tiny_ragged", "tiny_same1", "base_small"])
def test_oracle_reproduces_golden(name):
    """Golden vectors (oracle/make_golden.py) are recomputed bit-for-bit-close on this machine."""
    from oracle.make_golden import CASES
    shape, lens, S, alens, same = CASES[name]
    z = np.load(os.path.join(GOLD, name + ".npz"))
    sd = O.synthetic_state_dict(shape)
    assert float(sum(v.double().sum() for v in sd.values())) == pytest.approx(float(z["weight_checksum"]), rel=1e-9)
    ids, mask = torch.from_numpy(z["ids"]), torch.from_numpy(z["mask"])
    a_ids, a_mask = torch.from_numpy(z["anchor_ids"]), torch.from_numpy(z["anchor_mask"])
    anchors = [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(a_ids.shape[0])]
    with torch.no_grad():
        bank = O.build_bank(sd, anchors, shape)
        out = O.memory_forward(sd, ids, mask, torch.zeros_like(ids), bank, int(z["same_idx"]), shape)
    assert float((bank - torch.from_numpy(z["bank"])).abs().max()) < 1e-5
    assert float((out["logits"] - torch.from_numpy(z["logits"])).abs().max()) < 1e-5
    assert out["best_idx"].tolist() == z["best_idx"].tolist()


def test_single_head_oracle():
    shape = O.BERT_TINY
    sd = O.synthetic_state_dict(shape, model="single")
    ids, mask, tids = O.synthetic_ids(3, 12, lens=[12, 5, 9], vocab_size=shape.vocab_size)
    with torch.no_grad():
        out = O.single_forward(sd, ids, mask, tids, shape)
    assert out["probs"].shape == (3, 2) and torch.allclose(out["probs"].sum(-1), torch.ones(3))
