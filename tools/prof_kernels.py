"""Tiny harness for `ncu --set full`: runs ONE kernel class at the C2 shape a few times.
    ncu --set full --clock-control none --import-source on -k regex:<pat> -s 2 -c 1 -o gpurun_out/<name> \
        python tools/prof_kernels.py <gemm_qkv|gemm_ffn_up|ln_attn_out|ln_ffn_down|attention|layernorm|pool_match>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N  # noqa: E402

which = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
M, H, I, B, S = 32768, 768, 3072, 64, 512
torch.manual_seed(0)
dev = "cuda"
if which.startswith("gemm"):
    Nn, K, epi = {"gemm_qkv": (2304, 768, 0), "gemm_attn_out": (768, 768, 2), "gemm_ffn_up": (3072, 768, 1),
                  "gemm_ffn_down": (768, 3072, 2)}[which]
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(Nn, K, device=dev) * 0.05).half()
    bias = torch.randn(Nn, device=dev)
    resid = torch.randn(M, Nn, device=dev)
    out = torch.empty(M, Nn, device=dev, dtype=torch.float32 if epi == 2 else torch.float16)
    for _ in range(iters):
        N.gemm_f16(a, w, bias, epi, resid=resid, out=out)
elif which.startswith("ln_"):
    K = {"ln_attn_out": 768, "ln_ffn_down": 3072}[which]
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(768, K, device=dev) * 0.05).half()
    bias = torch.randn(768, device=dev); gamma = torch.ones(768, device=dev); beta = torch.zeros(768, device=dev)
    resid = torch.randn(M, 768, device=dev)
    for _ in range(iters):
        N.gemm_ln_f16(a, w, bias, resid, gamma, beta, inplace=True)
elif which == "attention":
    qkv = torch.randn(M, 3 * H, device=dev).half()
    lens = torch.full((B,), S, dtype=torch.int32, device=dev)
    for _ in range(iters):
        N.attention_f16(qkv, lens, B, S, H)
elif which == "layernorm":
    y = torch.randn(M, H, device=dev)
    g, b = torch.randn(H, device=dev), torch.randn(H, device=dev)
    for _ in range(iters):
        N.layernorm(y, g, b)
elif which == "pool_match":
    G, D = int(os.environ.get("PM_G", 129)), 512
    Bq = int(os.environ.get("PM_B", 64))
    cls = torch.randn(Bq, H, device=dev)
    wp, bp = torch.randn(H, H, device=dev) * 0.03, torch.randn(H, device=dev) * 0.02
    wh, bh = torch.randn(D, H, device=dev) * 0.03, torch.randn(D, device=dev) * 0.02
    wproj = torch.randn(2, 3 * D, device=dev) * 0.03
    bank = torch.relu(torch.randn(G, D, device=dev))
    vterm = N.bank_prepare(bank, wproj)
    for _ in range(iters):
        N.pool_match(cls, H, Bq, wp, bp, wh, bh, wproj, bank, vterm)
torch.cuda.synchronize()
print("done", which)
