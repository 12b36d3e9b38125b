"""GPU bring-up probe: runs every kernel against a torch / oracle reference, each group in its own
subprocess under a timeout (a trapped or hung kernel must not take the other groups down).
Usage on the GPU box:   python tools/gpu_probe.py [group ...]     (writes gpurun_out/probe_<group>.log)
"""
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GROUPS = ["gemm", "gemm_ln", "attention", "rowwise", "poolmatch", "encoder_tiny", "encoder_base", "model", "precise"]


def _report(name, got, ref, tol):
    import torch
    d = (got.float() - ref.float()).abs()
    bad = bool(torch.isnan(got.float()).any()) or float(d.max()) > tol
    print(f"  {'FAIL' if bad else 'ok  '} {name}: max|d|={float(d.max()):.3e} mean|d|={float(d.mean()):.3e} "
          f"ref_absmax={float(ref.abs().max()):.3e} tol={tol:g}", flush=True)
    return not bad


def g_gemm():
    import torch
    from memvul_b200 import native as N
    ok = True
    torch.manual_seed(0)
    dev = "cuda"
    for (M, Nn, K) in [(128, 128, 64), (128, 256, 128), (300, 384, 128), (512, 768, 768), (1000, 2304, 768),
                       (20000, 768, 768), (20000, 2304, 768), (4096, 3072, 768), (4096, 768, 3072), (32768, 768, 3072)]:
        a = (torch.randn(M, K, device=dev) * 1.0).half()
        w = (torch.randn(Nn, K, device=dev) * 0.05).half()
        bias = torch.randn(Nn, device=dev)
        resid = torch.randn(M, Nn, device=dev)
        ref = a.float() @ w.float().T + bias
        for epi, name in [(N.EPI_BIAS_F16, "bias"), (N.EPI_BIAS_GELU_F16, "gelu"), (N.EPI_BIAS_RESID_F32, "resid")]:
            if epi == N.EPI_BIAS_GELU_F16:
                r = torch.nn.functional.gelu(ref)
            elif epi == N.EPI_BIAS_RESID_F32:
                r = ref + resid
            else:
                r = ref
            out = N.gemm_f16(a, w, bias, epi, resid=resid if epi == N.EPI_BIAS_RESID_F32 else None)
            torch.cuda.synchronize()
            tol = 2e-3 * max(1.0, float(r.abs().max())) if epi != N.EPI_BIAS_RESID_F32 else 1e-3
            ok &= _report(f"gemm M={M} N={Nn} K={K} {name}", out, r, tol)
    # timing of the bert-base shapes at M=32768
    M = 32768
    for (Nn, K, epi) in [(2304, 768, 0), (768, 768, 2), (3072, 768, 1), (768, 3072, 2)]:
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(Nn, K, device=dev) * 0.05).half()
        bias = torch.randn(Nn, device=dev)
        resid = torch.randn(M, Nn, device=dev)
        out = torch.empty(M, Nn, device=dev, dtype=torch.float32 if epi == 2 else torch.float16)
        for _ in range(3):
            N.gemm_f16(a, w, bias, epi, resid=resid, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            N.gemm_f16(a, w, bias, epi, resid=resid, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  time gemm M={M} N={Nn} K={K} epi={epi}: {ms*1e3:.1f} us  {2.0*M*Nn*K/ms/1e9:.1f} TFLOP/s", flush=True)
    return ok


def g_gemm_ln():
    import torch
    from memvul_b200 import native as N
    ok = True
    torch.manual_seed(5)
    dev = "cuda"
    for (M, K) in [(256, 768), (1000, 768), (4096, 3072), (20001, 768), (32768, 768), (32768, 3072)]:
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(768, K, device=dev) * 0.05).half()
        bias = torch.randn(768, device=dev)
        resid = torch.randn(M, 768, device=dev) * 2 + 0.3
        gamma = 1 + 0.1 * torch.randn(768, device=dev)
        beta = 0.1 * torch.randn(768, device=dev)
        ref = torch.nn.functional.layer_norm(a.float() @ w.float().T + bias + resid, (768,), gamma, beta, 1e-12)
        x32, x16 = N.gemm_ln_f16(a, w, bias, resid, gamma, beta)
        torch.cuda.synchronize()
        ok &= _report(f"gemm_ln M={M} K={K} fp32", x32, ref, 2e-4)
        ok &= _report(f"gemm_ln M={M} K={K} fp16", x16, ref, 4e-3)
        buf = resid.clone()
        y32, _ = N.gemm_ln_f16(a, w, bias, buf, gamma, beta, inplace=True)
        torch.cuda.synchronize()
        ok &= _report(f"gemm_ln M={M} K={K} in-place", y32, ref, 2e-4)
    M = 32768
    for K in (768, 3072):
        a = torch.randn(M, K, device=dev).half()
        w = (torch.randn(768, K, device=dev) * 0.05).half()
        bias = torch.randn(768, device=dev); gamma = torch.ones(768, device=dev); beta = torch.zeros(768, device=dev)
        resid = torch.randn(M, 768, device=dev)
        for _ in range(3):
            N.gemm_ln_f16(a, w, bias, resid, gamma, beta, inplace=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            N.gemm_ln_f16(a, w, bias, resid, gamma, beta, inplace=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  time gemm_ln M={M} K={K}: {ms*1e3:.1f} us  {2.0*M*768*K/ms/1e9:.1f} TFLOP/s", flush=True)
    return ok


def _attn_ref(qkv, lens, B, S, H):
    import torch
    nH = H // 64
    q, k, v = qkv.float().view(B, S, 3, nH, 64).permute(2, 0, 3, 1, 4)
    sc = q @ k.transpose(-1, -2) / 8.0
    mask = torch.arange(S, device=qkv.device)[None, :] < lens[:, None]
    sc = sc + (1.0 - mask.float())[:, None, None, :] * -10000.0
    return (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, H), mask


def g_attention():
    import torch
    from memvul_b200 import native as N
    ok = True
    torch.manual_seed(1)
    for (B, S, H, lens) in [(1, 128, 128, [128]), (2, 128, 128, [128, 77]), (2, 256, 128, [256, 130]),
                            (3, 512, 768, [512, 300, 5]), (2, 200, 768, [200, 129]), (4, 64, 128, [64, 2, 33, 17]),
                            (8, 512, 768, [512] * 8)]:
        qkv = (torch.randn(B * S, 3 * H, device="cuda") * 1.5).half()
        lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
        ctx = N.attention_f16(qkv, lens_t, B, S, H)
        torch.cuda.synchronize()
        ref, mask = _attn_ref(qkv, lens_t, B, S, H)
        valid = mask.reshape(-1)
        ok &= _report(f"attention B={B} S={S} H={H} lens={lens if len(lens) < 5 else 'full'}", ctx[valid], ref[valid], 4e-3)
    B, S, H = 64, 512, 768
    qkv = torch.randn(B * S, 3 * H, device="cuda").half()
    lens_t = torch.full((B,), S, dtype=torch.int32, device="cuda")
    for _ in range(3):
        N.attention_f16(qkv, lens_t, B, S, H)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        N.attention_f16(qkv, lens_t, B, S, H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"  time attention B={B} S={S}: {ms*1e3:.1f} us  {4.0*B*12*S*S*64/ms/1e9:.1f} TFLOP/s", flush=True)
    return ok


def g_rowwise():
    import torch
    from memvul_b200 import native as N
    from oracle import memvul_oracle as O
    ok = True
    torch.manual_seed(2)
    for H in (768, 128):
        y = torch.randn(1000, H, device="cuda") * 3 + 0.5
        g = torch.randn(H, device="cuda")
        b = torch.randn(H, device="cuda")
        x32, x16 = N.layernorm(y, g, b)
        ref = torch.nn.functional.layer_norm(y, (H,), g, b, 1e-12)
        ok &= _report(f"layernorm H={H} fp32", x32, ref, 2e-5)
        ok &= _report(f"layernorm H={H} fp16", x16, ref, 4e-3)
    for shape in (O.BERT_TINY, O.BERT_BASE):
        sd = O.synthetic_state_dict(shape)
        w = N.PackedBert(sd, O.EMB, torch.device("cuda"))
        ids, mask, tids = O.synthetic_ids(3, 40, lens=[40, 7, 22], vocab_size=shape.vocab_size)
        tids[0, 5:] = 1
        x32, _ = N.embed_layernorm(w, ids.cuda(), tids.cuda())
        e = O.EMB + "embeddings."
        ref = O._ln(sd[e + "word_embeddings.weight"][ids] + sd[e + "position_embeddings.weight"][:40][None]
                    + sd[e + "token_type_embeddings.weight"][tids], sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], 1e-12)
        ok &= _report(f"embed H={shape.hidden}", x32.cpu().view(3, 40, -1), ref, 2e-5)
    m = torch.zeros(5, 50, dtype=torch.bool, device="cuda")
    for i, l in enumerate([50, 1, 17, 32, 49]):
        m[i, :l] = True
    lens, bad = N.mask_to_lens(m)
    print("  mask_to_lens", lens.tolist(), "bad", int(bad), flush=True)
    ok &= lens.tolist() == [50, 1, 17, 32, 49] and int(bad) == 0
    m[2, 30] = True
    _, bad = N.mask_to_lens(m)
    ok &= int(bad) == 1
    return ok


def g_poolmatch():
    import torch
    from memvul_b200 import native as N
    from oracle import memvul_oracle as O
    ok = True
    torch.manual_seed(3)
    for (B, G, H, D, same) in [(4, 129, 768, 512, 0), (64, 129, 768, 512, 1), (7, 1, 768, 512, 0), (5, 6, 128, 64, 0),
                               (33, 1000, 768, 512, 0), (256, 4096, 768, 512, 0)]:
        cls = torch.randn(B, 3, H)
        wp, bp = torch.randn(H, H) * 0.03, torch.randn(H) * 0.02
        wh, bh = torch.randn(D, H) * 0.03, torch.randn(D) * 0.02
        wproj = torch.randn(2, 3 * D) * 0.03
        bank = torch.relu(torch.randn(G, D) * 0.4)
        lin = torch.nn.functional.linear
        u_ref = torch.relu(lin(torch.tanh(lin(cls[:, 0], wp, bp)), wh, bh))
        ref = O.match(u_ref, bank, wproj, same)
        c = lambda t: t.cuda().contiguous()
        clsd, bankd, wprojd = c(cls), c(bank), c(wproj)
        vterm = N.bank_prepare(bankd, wprojd)
        out = N.pool_match(clsd, 3 * H, B, c(wp), c(bp), c(wh), c(bh), wprojd, bankd, vterm, same_idx=same)
        torch.cuda.synchronize()
        tag = f"B={B} G={G} H={H} D={D} same={same}"
        ok &= _report(f"poolmatch u {tag}", out["u"].cpu(), u_ref, 2e-5)
        ok &= _report(f"poolmatch logits {tag}", out["logits"].cpu(), ref["logits"], 5e-5)
        ok &= _report(f"poolmatch probs {tag}", out["probs"].cpu(), ref["p"], 2e-5)
        ok &= _report(f"poolmatch best_probs {tag}", out["best_probs"].cpu(), ref["probs"], 2e-5)
        # argmax identical except where the top-2 p_same gap is below fp32 noise
        idx = out["best_idx"].cpu().long()
        ps = ref["p"][:, :, same]
        gap = ps.max(1).values - ps[torch.arange(B), idx]
        nmis = int((idx != ref["best_idx"]).sum())
        print(f"  {'ok  ' if float(gap.max()) < 1e-6 else 'FAIL'} poolmatch argmax {tag}: mismatches={nmis} max gap={float(gap.max()):.2e}", flush=True)
        ok &= float(gap.max()) < 1e-6
    # exact-tie test: duplicate anchors -> lowest index must win
    B, G, H, D = 3, 8, 768, 512
    u = torch.relu(torch.randn(B, D)).cuda()
    bank = torch.relu(torch.randn(1, D)).repeat(G, 1).cuda()
    wproj = (torch.randn(2, 3 * D) * 0.03).cuda()
    vterm = N.bank_prepare(bank, wproj)
    dummy = torch.zeros(H, H, device="cuda")
    out = N.pool_match(None, 0, B, dummy, None, torch.zeros(D, H, device="cuda"), None, wproj, bank, vterm,
                       phase_mask=N.PM_UTERM | N.PM_MATCH | N.PM_FINAL, u=u)
    torch.cuda.synchronize()
    print("  tie best_idx", out["best_idx"].tolist(), flush=True)
    ok &= out["best_idx"].tolist() == [0, 0, 0]
    return ok


def _encoder(shape, B, S, lens, tol):
    import torch
    from memvul_b200 import native as N
    from oracle import memvul_oracle as O
    sd = O.synthetic_state_dict(shape)
    w = N.PackedBert(sd, O.EMB, torch.device("cuda"))
    ids, mask, tids = O.synthetic_ids(B, S, lens=lens, vocab_size=shape.vocab_size)
    lens_t, bad = N.mask_to_lens(mask.cuda())
    hid = N.encoder_forward(w, ids.cuda(), lens_t)
    torch.cuda.synchronize()
    t0 = time.time()
    ref = O.bert_encoder(sd, ids, mask.float(), None, shape)
    print(f"  oracle encoder took {time.time()-t0:.1f}s", flush=True)
    ok = _report(f"encoder H={shape.hidden} L={shape.layers} B={B} S={S} CLS", hid[:, 0].cpu(), ref[:, 0], tol)
    ok &= _report(f"encoder all valid rows", hid.cpu()[mask], ref[mask], tol * 2)
    return ok


def g_encoder_tiny():
    from oracle import memvul_oracle as O
    return _encoder(O.BERT_TINY, 4, 128, [128, 60, 2, 100], 1e-2) & _encoder(O.BERT_TINY, 3, 200, [200, 131, 17], 1e-2)


def g_encoder_base():
    from oracle import memvul_oracle as O
    return _encoder(O.BERT_BASE, 4, 128, [128, 100, 64, 17], 1.5e-2) & _encoder(O.BERT_BASE, 2, 512, [512, 300], 1.5e-2)


def g_precise():
    """Accuracy mode (MEMVUL_ENC_PRECISE) building blocks and the tiny encoder, padded and packed; also the default packed
    encoder (speculative-softmax attention, third-residual-buffer GEMM+LN at M >= 256)."""
    import torch
    from memvul_b200 import native as N
    from oracle import memvul_oracle as O
    ok = True
    torch.manual_seed(3)
    x = torch.randn(300, 768, device="cuda")
    sp = N.split3_f16(x)
    ok &= _report("split3 hi+lo", sp[:, :768].float() + sp[:, 768:1536].float(), x, 1e-6)
    w = torch.randn(768, 768, device="cuda") * 0.05
    bias = torch.randn(768, device="cuda")
    out = N.gemm_f16(sp, N.split3_weight(w), bias, N.EPI_BIAS_F32)
    ok &= _report("split GEMM 300x768x768 (fp32 out, no residual)", out, x @ w.T + bias, 1e-4)
    qkv = torch.randn(2 * 200, 3 * 128, device="cuda")
    lens = torch.tensor([200, 77], dtype=torch.int32, device="cuda")
    ctx = N.attention_f32(qkv, lens, 2, 200, 128)
    ref, valid = _attn_ref(qkv, lens, 2, 200, 128)
    valid = valid.reshape(-1)
    ok &= _report("attention_f32 B=2 S=200", ctx[valid], ref[valid], 1e-4)
    sd = O.synthetic_state_dict(O.BERT_TINY)
    ids, mask, tids = O.synthetic_ids(3, 300, lens=[300, 131, 17], vocab_size=O.BERT_TINY.vocab_size)
    refh = O.bert_encoder(sd, ids, mask.float(), None, O.BERT_TINY)
    for precise in (True, False):
        wts = N.PackedBert(sd, O.EMB, torch.device("cuda"), precise=precise)
        for packed in (False, True):
            if packed:
                lens_t, rs, bad = N.mask_to_lens(mask.cuda(), with_row_start=True)
            else:
                (lens_t, bad), rs = N.mask_to_lens(mask.cuda()), None
            hid = N.encoder_forward(wts, ids.cuda(), lens_t, row_start=rs, bad=bad)
            ok &= _report(f"encoder tiny precise={precise} packed={packed}", hid.cpu()[mask], refh[mask], 2e-5 if precise else 2e-2)
    return ok


def g_model():
    print("  (model group filled in once memvul_b200.model_memory exists)")
    try:
        import tools.gpu_probe_model as pm      # noqa
        return pm.run()
    except ImportError:
        return True


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        import torch
        print(f"[{sys.argv[2]}] device {torch.cuda.get_device_name(0)}", flush=True)
        ok = globals()["g_" + sys.argv[2]]()
        print(f"[{sys.argv[2]}] {'PASS' if ok else 'FAIL'}", flush=True)
        sys.exit(0 if ok else 1)
    groups = sys.argv[1:] or GROUPS
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    summary = []
    for g in groups:
        t0 = time.time()
        log = os.path.join(ROOT, "gpurun_out", f"probe_{g}.log")
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", g], stdout=f,
                                   stderr=subprocess.STDOUT, timeout=420, cwd=ROOT)
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        summary.append(f"{g}: rc={rc} ({time.time()-t0:.0f}s)")
        print(summary[-1], flush=True)
        print(open(log).read()[-3000:], flush=True)
    print("SUMMARY " + " | ".join(summary))


if __name__ == "__main__":
    main()
