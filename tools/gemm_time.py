"""Times the four bert-base GEMM shapes at M=32768 (C2) back to back; env knobs select kernel variants."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N
M = 32768
dev = "cuda"
for (Nn, K, epi) in [(2304, 768, 0), (768, 768, 2), (3072, 768, 1), (768, 3072, 2)]:
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(Nn, K, device=dev) * 0.05).half()
    bias = torch.randn(Nn, device=dev); resid = torch.randn(M, Nn, device=dev)
    out = torch.empty(M, Nn, device=dev, dtype=torch.float32 if epi == 2 else torch.float16)
    for _ in range(3): N.gemm_f16(a, w, bias, epi, resid=resid, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): N.gemm_f16(a, w, bias, epi, resid=resid, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"  N={Nn} K={K} epi={epi}: {ms*1e3:.1f} us  {2.0*M*Nn*K/ms/1e9:.1f} TFLOP/s", flush=True)
for K in (768, 3072):                                     # fused GEMM + residual + LayerNorm (N = 768), in place like the encoder
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(768, K, device=dev) * 0.05).half()
    bias = torch.randn(768, device=dev); resid = torch.randn(M, 768, device=dev)
    g = torch.rand(768, device=dev) + 0.5; b = torch.randn(768, device=dev)
    for _ in range(3): N.gemm_ln_f16(a, w, bias, resid, g, b, inplace=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): N.gemm_ln_f16(a, w, bias, resid, g, b, inplace=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"  gemm_ln K={K}: {ms*1e3:.1f} us  {2.0*M*768*K/ms/1e9:.1f} TFLOP/s", flush=True)
