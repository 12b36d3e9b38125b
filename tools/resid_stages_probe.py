"""Experiment (r02o): the residual-epilogue GEMM (N=768) with four vs five A/B stages (MEMVUL_LIB_PATH selects the build)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N
M = 32768
for (Nn, K) in [(768, 3072), (768, 768)]:
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(Nn, K, device="cuda") * 0.05).half()
    bias = torch.randn(Nn, device="cuda"); resid = torch.randn(M, Nn, device="cuda")
    out = torch.empty(M, Nn, device="cuda")
    N.gemm_f16(a, w, bias, 2, resid=resid, out=out)
    ref = a[:512].float() @ w.float().T + bias + resid[:512]
    err = float((out[:512] - ref).abs().max())
    for _ in range(3): N.gemm_f16(a, w, bias, 2, resid=resid, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): N.gemm_f16(a, w, bias, 2, resid=resid, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{os.environ.get('MEMVUL_LIB_PATH', 'default')}: N={Nn} K={K} epi=2: {ms*1e3:.1f} us  {2.0*M*Nn*K/ms/1e9:.1f} TFLOP/s  max err {err:.2e}", flush=True)
