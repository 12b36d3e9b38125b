"""GPU parity of each sm_100a kernel, called through the C ABI, against a plain fp32 PyTorch statement of the
same op on the same (fp16-rounded) inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def N():
    from memvul_b200 import native
    native.build()
    return native


@pytest.mark.parametrize("M,Nn,K", [(128, 128, 64), (1, 128, 64), (300, 384, 128), (512, 768, 768), (1000, 2304, 768),
                                    (4096, 3072, 768), (4096, 768, 3072), (20000, 768, 768)])
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_gemm_tcgen05(N, M, Nn, K, epi):
    """SURVEY 2.2 K2/K4/K5/K6.  Tolerance: fp16 output rounding (2^-11 relative) for epilogues 0/1; fp32
    accumulation-order noise for the fp32 residual epilogue."""
    torch.manual_seed(M + Nn + K + epi)
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(Nn, K, device="cuda") * 0.05).half()
    bias = torch.randn(Nn, device="cuda")
    resid = torch.randn(M, Nn, device="cuda")
    ref = a.float() @ w.float().T + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        ref = ref + resid
    out = N.gemm_f16(a, w, bias, epi, resid=resid if epi == 2 else None)
    tol = 1e-3 if epi == 2 else 1.5e-3 * max(1.0, float(ref.abs().max()))
    assert not torch.isnan(out.float()).any()
    assert float((out.float() - ref).abs().max()) < tol
    if epi == 2:                                   # in-place residual (how the encoder uses it)
        buf = resid.clone()
        N.gemm_f16(a, w, bias, 2, resid=buf, out=buf)
        assert float((buf - ref).abs().max()) < 1e-3


@pytest.mark.parametrize("M,K", [(256, 768), (1000, 768), (4096, 3072), (20001, 768)])
def test_gemm_residual_layernorm_fused(N, M, K):
    """SURVEY 2.2 K4/K6 in one kernel (six-CTA clusters): LayerNorm(A W^T + bias + resid) -> fp32 and fp16, also in place."""
    torch.manual_seed(M + K)
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(768, K, device="cuda") * 0.05).half()
    bias = torch.randn(768, device="cuda")
    resid = torch.randn(M, 768, device="cuda") * 2 + 0.3
    gamma = 1 + 0.1 * torch.randn(768, device="cuda")
    beta = 0.1 * torch.randn(768, device="cuda")
    ref = torch.nn.functional.layer_norm(a.float() @ w.float().T + bias + resid, (768,), gamma, beta, 1e-12)
    x32, x16 = N.gemm_ln_f16(a, w, bias, resid, gamma, beta)
    assert float((x32 - ref).abs().max()) < 2e-4
    assert float((x16.float() - ref).abs().max()) < 4e-3
    buf = resid.clone()
    y32, y16 = N.gemm_ln_f16(a, w, bias, buf, gamma, beta, inplace=True)
    assert torch.equal(y32, x32) and torch.equal(y16, x16)          # deterministic, in place == out of place
    with pytest.raises(ValueError):
        N.gemm_ln_f16(a[:128].contiguous(), w, bias, resid[:128].contiguous(), gamma, beta)


def _attn_ref(qkv, lens, B, S, H):
    nH = H // 64
    q, k, v = qkv.float().view(B, S, 3, nH, 64).permute(2, 0, 3, 1, 4)
    mask = torch.arange(S, device=qkv.device)[None, :] < lens[:, None]
    sc = q @ k.transpose(-1, -2) / 8.0 + (1.0 - mask.float())[:, None, None, :] * -10000.0     # transformers 4.1.0 mask
    return (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, H), mask.reshape(-1)


@pytest.mark.parametrize("B,S,H,lens", [(1, 128, 128, [128]), (2, 128, 128, [128, 77]), (2, 256, 128, [256, 130]),
                                        (3, 512, 768, [512, 300, 5]), (2, 200, 768, [200, 129]),
                                        (4, 64, 128, [64, 2, 33, 17]), (2, 1, 128, [1, 1]), (2, 511, 128, [511, 384])])
def test_attention_tcgen05(N, B, S, H, lens):
    """softmax(QK^T/8 + (1-m)(-1e4)) V on valid query rows (padded rows are unspecified)."""
    torch.manual_seed(S + H)
    qkv = (torch.randn(B * S, 3 * H, device="cuda") * 1.5).half()
    lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
    ctx = N.attention_f16(qkv, lens_t, B, S, H)
    ref, valid = _attn_ref(qkv, lens_t, B, S, H)
    assert not torch.isnan(ctx.float()).any()
    assert float((ctx.float() - ref)[valid].abs().max()) < 4e-3


@pytest.mark.parametrize("H", [768, 128])
def test_layernorm_rows(N, H):
    torch.manual_seed(H)
    y = torch.randn(777, H, device="cuda") * 3 + 0.5
    g, b = torch.randn(H, device="cuda"), torch.randn(H, device="cuda")
    x32, x16 = N.layernorm(y, g, b)
    ref = torch.nn.functional.layer_norm(y, (H,), g, b, 1e-12)
    assert float((x32 - ref).abs().max()) < 2e-5
    assert float((x16.float() - ref).abs().max()) < 4e-3 * max(1.0, float(ref.abs().max()) / 4)


def test_embed_layernorm_and_mask_to_lens(N):
    from memvul_b200.synthetic import BERT_TINY, EMB, synthetic_ids, synthetic_state_dict
    sd = synthetic_state_dict(BERT_TINY)
    w = N.PackedBert(sd, EMB, torch.device("cuda"))
    ids, mask, tids = synthetic_ids(3, 40, lens=[40, 7, 22], vocab_size=1024)
    tids[0, 5:] = 1
    x32, x16 = N.embed_layernorm(w, ids.cuda(), tids.cuda())
    e = EMB + "embeddings."
    x = sd[e + "word_embeddings.weight"][ids] + sd[e + "position_embeddings.weight"][:40][None] + sd[e + "token_type_embeddings.weight"][tids]
    ref = torch.nn.functional.layer_norm(x, (128,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], 1e-12)
    assert float((x32.cpu().view(3, 40, 128) - ref).abs().max()) < 2e-5
    lens, bad = N.mask_to_lens(mask.cuda())
    assert lens.tolist() == [40, 7, 22] and int(bad) == 0
    m2 = mask.clone(); m2[1, 20] = True
    assert int(N.mask_to_lens(m2.cuda())[1]) == 1
    m3 = mask.clone(); m3[2] = False
    assert int(N.mask_to_lens(m3.cuda())[1]) == 1


@pytest.mark.parametrize("B,G,H,D,same", [(4, 129, 768, 512, 0), (64, 129, 768, 512, 1), (7, 1, 768, 512, 0),
                                          (5, 6, 128, 64, 0), (33, 1000, 768, 512, 0), (256, 2048, 768, 512, 1)])
def test_pool_match_fused(N, B, G, H, D, same):
    """SURVEY 2.2 K7-K10 against the literal concat/Linear/softmax/argmax statement of model_memory.py:135-147."""
    from oracle import memvul_oracle as O
    torch.manual_seed(B * 7 + G)
    cls = torch.randn(B, 3, H)
    wp, bp = torch.randn(H, H) * 0.03, torch.randn(H) * 0.02
    wh, bh = torch.randn(D, H) * 0.03, torch.randn(D) * 0.02
    wproj = torch.randn(2, 3 * D) * 0.03
    bank = torch.relu(torch.randn(G, D) * 0.4)
    lin = torch.nn.functional.linear
    u_ref = torch.relu(lin(torch.tanh(lin(cls[:, 0], wp, bp)), wh, bh))
    ref = O.match(u_ref, bank, wproj, same)
    c = lambda t: t.cuda().contiguous()
    bankd, wprojd = c(bank), c(wproj)
    out = N.pool_match(c(cls), 3 * H, B, c(wp), c(bp), c(wh), c(bh), wprojd, bankd, N.bank_prepare(bankd, wprojd), same_idx=same)
    assert float((out["u"].cpu() - u_ref).abs().max()) < 2e-5
    assert float((out["logits"].cpu() - ref["logits"]).abs().max()) < 5e-5
    assert float((out["probs"].cpu() - ref["p"]).abs().max()) < 2e-5
    assert float((out["best_probs"].cpu() - ref["probs"]).abs().max()) < 2e-5
    idx = out["best_idx"].cpu().long()
    ps = ref["p"][:, :, same]
    assert float((ps.max(1).values - ps[torch.arange(B), idx]).max()) < 1e-6     # identical unless an fp32-noise tie
    # the kernel's own outputs are self-consistent bit for bit: best = first maximum of its probs
    pk = out["probs"][:, :, same]
    assert torch.equal(out["best_idx"].long(), pk.argmax(1))
    assert torch.equal(out["best_probs"], out["probs"][torch.arange(B, device="cuda"), out["best_idx"].long()])
    # phase-by-phase launches give the same bits as the fused cooperative launch
    st = N.pool_match(c(cls), 3 * H, B, c(wp), c(bp), c(wh), c(bh), phase_mask=N.PM_POOL | N.PM_HEADER)
    assert torch.equal(st["u"], out["u"])


def test_pool_match_ties_pick_lowest_index(N):
    B, G, H, D = 3, 8, 768, 512
    torch.manual_seed(0)
    u = torch.relu(torch.randn(B, D)).cuda()
    bank = torch.relu(torch.randn(1, D)).repeat(G, 1).cuda()
    wproj = (torch.randn(2, 3 * D) * 0.03).cuda()
    out = N.pool_match(None, 0, B, torch.zeros(H, H, device="cuda"), None, torch.zeros(D, H, device="cuda"), None, wproj,
                       bank, N.bank_prepare(bank, wproj), phase_mask=N.PM_UTERM | N.PM_MATCH | N.PM_FINAL, u=u)
    assert out["best_idx"].tolist() == [0, 0, 0]


def test_single_head(N):
    torch.manual_seed(1)
    f = torch.randn(9, 512, device="cuda")
    w = torch.randn(2, 512, device="cuda") * 0.05
    logits, probs = N.single_head(f, w)
    assert float((logits - f @ w.T).abs().max()) < 1e-5 and float((probs - torch.softmax(f @ w.T, -1)).abs().max()) < 1e-6


def test_errors_are_loud(N):
    a = torch.zeros(128, 60, device="cuda", dtype=torch.float16)
    with pytest.raises(ValueError):
        N.gemm_f16(a, torch.zeros(128, 60, device="cuda", dtype=torch.float16), torch.zeros(128, device="cuda"), 0)
    with pytest.raises(ValueError):
        N.attention_f16(torch.zeros(1024, 384, device="cuda", dtype=torch.float16), torch.ones(1, dtype=torch.int32, device="cuda"), 1, 1024, 128)
    with pytest.raises(N.NativeError):
        N.gemm_f16(torch.zeros(128, 64, dtype=torch.float16), torch.zeros(128, 64, dtype=torch.float16), torch.zeros(128), 0)
