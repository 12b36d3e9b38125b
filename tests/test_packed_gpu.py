"""Packed (token-major, var-len) execution, deferred error flags, cache invalidation, device guard.

The reference pads every batch to its longest member (config_memory.json:50-57) and pays for the padding in every GEMM;
MEMVUL_ENC_PACKED computes only the valid tokens.  Row arithmetic is unchanged, so the packed and the padded execution
must agree bit for bit on every valid row."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def N():
    from memvul_b200 import native
    native.build()
    return native


def _attention_ref(qkv, lens, B, S, H, row_start=None):
    """fp32 softmax(QK^T/8 + key mask)V per head on the same fp16-rounded inputs; rows past len are skipped."""
    nh = H // 64
    q, k, v = qkv.float().split(H, dim=1)
    out = torch.zeros(qkv.shape[0], H)
    for b in range(B):
        r0 = int(row_start[b]) if row_start is not None else b * S
        L = int(lens[b])
        for h in range(nh):
            sl = slice(h * 64, h * 64 + 64)
            s = (q[r0:r0 + L, sl] @ k[r0:r0 + L, sl].T) / 8.0
            out[r0:r0 + L, sl] = torch.softmax(s, -1) @ v[r0:r0 + L, sl]
    return out


@pytest.mark.parametrize("lens", [[512, 1, 130, 64, 300], [128, 128], [5], [257, 511, 129, 63, 200, 31]])
def test_attention_packed_layout(N, lens):
    H, B, S = 128, len(lens), max(lens)
    torch.manual_seed(sum(lens))
    T = sum(lens)
    rs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    qkv = (torch.randn(B * S, 3 * H) * 0.7).half()                  # B*S rows = the upper bound the TMA maps are sized for
    lens_t = torch.tensor(lens, dtype=torch.int32)
    ctx = N.attention_f16(qkv.cuda(), lens_t.cuda(), B, S, H, row_start=rs.cuda()).cpu().float()
    ref = _attention_ref(qkv, lens, B, S, H, rs)
    assert float((ctx[:T] - ref[:T]).abs().max()) < 4e-3
    if T < B * S:
        assert float(ctx[T:].abs().max()) == 0.0                     # nothing is written past the last token


@pytest.mark.parametrize("shape_name,lens,S", [("tiny", [200, 130, 7, 64, 129, 2], 200), ("base", [256, 100, 31, 255], 256),
                                               ("base", [512, 512], 512)])
def test_packed_encoder_equals_padded(N, shape_name, lens, S):
    from memvul_b200.synthetic import BERT_BASE, BERT_TINY, EMB, synthetic_ids, synthetic_state_dict
    shape = BERT_TINY if shape_name == "tiny" else BERT_BASE
    w = N.PackedBert(synthetic_state_dict(shape), EMB, torch.device("cuda"))
    ids, mask, _ = synthetic_ids(len(lens), S, lens=lens, vocab_size=shape.vocab_size)
    lens_t, row_start, bad = N.mask_to_lens(mask.cuda(), with_row_start=True)
    assert row_start.tolist() == [0] + torch.tensor(lens).cumsum(0).tolist() and int(bad) == 0
    full_pad = N.encoder_forward(w, ids.cuda(), lens_t)
    full_pack = N.encoder_forward(w, ids.cuda(), lens_t, row_start=row_start)
    m = mask.cuda()
    assert torch.equal(full_pack[m], full_pad[m])                    # same arithmetic per valid row: same bits
    assert float(full_pack[~m].abs().max() if (~m).any() else 0.0) == 0.0     # padded positions are zero-filled
    cls_pad = N.encoder_forward(w, ids.cuda(), lens_t, cls_only=True)
    cls_pack = N.encoder_forward(w, ids.cuda(), lens_t, cls_only=True, row_start=row_start)
    assert torch.equal(cls_pack[:, 0], cls_pad[:, 0])
    assert float((cls_pack[:, 0] - full_pad[:, 0]).abs().max()) < 1e-5


def test_out_of_range_ids_are_reported_not_silently_clamped(N):
    """torch.embedding raises on an out-of-range id and custom_PTM_embedder.py:205 on a type id; the kernel
    clamps but sets bit 1 of the deferred flag, which every host read turns into ValueError."""
    from memvul_b200.synthetic import BERT_TINY, build_memory_model, synthetic_ids
    model, _ = build_memory_model(BERT_TINY, device="cuda")
    ids, mask, tids = synthetic_ids(3, 16, lens=[16, 5, 9], vocab_size=1024)
    meta = [{"type": "golden", "instance": [{"label": f"CWE-{i}"}]} for i in range(3)]
    d = lambda i, m, t: {"tokens": {"token_ids": i.cuda(), "mask": m.cuda(), "type_ids": t.cuda()}}
    bad_ids = ids.clone(); bad_ids[1, 2] = 1024
    with torch.no_grad(), pytest.raises(ValueError, match="out of range"):
        model.forward_gold_instances(d(bad_ids, mask, tids), meta)
    bad_t = tids.clone(); bad_t[0, 3] = 2
    with torch.no_grad(), pytest.raises(ValueError, match="out of range"):
        model.forward_gold_instances(d(ids, mask, bad_t), meta)
    pad_garbage = ids.clone(); pad_garbage[1, 10] = 99999           # a PADDED position is never read: no error
    with torch.no_grad():
        model.forward_gold_instances(d(pad_garbage, mask, tids), meta)
        out = model(sample1=d(bad_ids, mask, tids), metadata=[{"type": "unlabel", "instance": [{"label": "neg", "Issue_Url": "u"}]}] * 3)
    with pytest.raises(ValueError, match="out of range"):
        out["probs"].tolist()


def test_bank_rebuild_with_same_shape_invalidates_the_cached_side_term():
    """a reset + rebuild with the same anchor count and an unchanged projector must not reuse the
    stale Wv.bank term (CustomValidation rebuilds the memory every epoch, callbacks.py:43-53)."""
    from memvul_b200.synthetic import BERT_TINY, build_memory_model, synthetic_ids
    sets = []
    for seed in (1, 2):
        a_ids, a_mask, _ = synthetic_ids(6, 32, lens=[32, 8, 20, 5, 32, 11], seed=seed, vocab_size=1024)
        sets.append((a_ids, a_mask))
    ids, mask, tids = synthetic_ids(4, 32, lens=[32, 9, 17, 30], seed=9, vocab_size=1024)
    d = lambda i, m: {"tokens": {"token_ids": i.cuda(), "mask": m.cuda(), "type_ids": torch.zeros_like(i).cuda()}}
    meta = [{"type": "golden", "instance": [{"label": f"CWE-{i}"}]} for i in range(6)]

    def run(model, which):
        model._golden_instances_embeddings = None                   # callbacks.py:48-49
        model._golden_instances_labels = None
        with torch.no_grad():
            model.forward_gold_instances(d(*sets[which]), meta)
            return model.match_batch(d(ids, mask))["logits"].clone()
    model, _ = build_memory_model(BERT_TINY, device="cuda")
    l_a = run(model, 0)
    l_b = run(model, 1)                                              # same G, same projector, different anchors
    fresh, _ = build_memory_model(BERT_TINY, device="cuda")
    l_b_fresh = run(fresh, 1)
    assert not torch.equal(l_a, l_b)
    assert torch.equal(l_b, l_b_fresh)


def test_model_on_a_second_device_while_device0_is_current():
    """``model.cuda(1)`` without ``set_device`` must run on device 1 (device guard + per-device
    kernel attributes)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from memvul_b200.synthetic import BERT_TINY, build_memory_model, synthetic_ids
    torch.cuda.set_device(0)
    m0, _ = build_memory_model(BERT_TINY, device="cuda:0")
    m1, _ = build_memory_model(BERT_TINY, device="cuda:1")
    a_ids, a_mask, _ = synthetic_ids(4, 40, lens=[40, 9, 22, 31], seed=2, vocab_size=1024)
    ids, mask, tids = synthetic_ids(3, 300, lens=[300, 120, 7], seed=3, vocab_size=1024)
    outs = []
    for m, dev in ((m0, "cuda:0"), (m1, "cuda:1")):
        to = lambda t: t.to(dev)
        with torch.no_grad():
            m.forward_gold_instances({"tokens": {"token_ids": to(a_ids), "mask": to(a_mask), "type_ids": to(torch.zeros_like(a_ids))}},
                                     [{"type": "golden", "instance": [{"label": f"CWE-{i}"}]} for i in range(4)])
            outs.append(m.match_batch({"tokens": {"token_ids": to(ids), "mask": to(mask), "type_ids": to(tids)}})["logits"].cpu())
    assert torch.equal(outs[0], outs[1])
