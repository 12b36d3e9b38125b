"""SM clock / board power while ONE kernel class runs back to back for ~1.5 s (is a kernel power-limited?).
    python tools/clock_probe.py <gemm_qkv|gemm_ffn_up|ln_attn_out|ln_ffn_down|attention>   (env knobs select variants)"""
import os, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N
import pynvml
which = sys.argv[1]
M, H, I, B, S = 32768, 768, 3072, 64, 512
dev = "cuda"
torch.manual_seed(0)
if which.startswith("gemm"):
    Nn, K, epi = {"gemm_qkv": (2304, 768, 0), "gemm_ffn_up": (3072, 768, 1)}[which]
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(Nn, K, device=dev) * 0.05).half()
    bias = torch.randn(Nn, device=dev); out = torch.empty(M, Nn, device=dev, dtype=torch.float16)
    run = lambda: N.gemm_f16(a, w, bias, epi, out=out)
    flops = 2.0 * M * Nn * K
elif which.startswith("ln_"):
    K = {"ln_attn_out": 768, "ln_ffn_down": 3072}[which]
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(768, K, device=dev) * 0.05).half()
    bias = torch.randn(768, device=dev); g = torch.ones(768, device=dev); b = torch.zeros(768, device=dev)
    resid = torch.randn(M, 768, device=dev)
    run = lambda: N.gemm_ln_f16(a, w, bias, resid, g, b, inplace=True)
    flops = 2.0 * M * 768 * K
else:
    qkv = torch.randn(M, 3 * H, device=dev).half(); lens = torch.full((B,), S, dtype=torch.int32, device=dev)
    run = lambda: N.attention_f16(qkv, lens, B, S, H)
    flops = 4.0 * B * 12 * S * S * 64
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.02)
for _ in range(50): run()
torch.cuda.synchronize()
t = threading.Thread(target=sampler); t.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 12000
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
stop = True; t.join()
ms = e0.elapsed_time(e1) / n
body = samples[len(samples) // 4:]
clk = sorted(c for c, _ in body); pw = sorted(p for _, p in body)
print(f"{which} [{' '.join(k + '=' + v for k, v in os.environ.items() if k.startswith('MEMVUL_'))}]: {ms*1e3:.1f} us/launch "
      f"{flops/ms/1e9:.0f} TF/s | SM clock median {clk[len(clk)//2]} MHz (min {clk[0]}, max {clk[-1]}) | power median {pw[len(pw)//2]:.0f} W (max {pw[-1]:.0f}) | {len(body)} samples")
