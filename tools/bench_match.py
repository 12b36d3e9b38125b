"""Anchor-match micro-benchmark (BASELINE.json metric part 2: "anchor-match HBM GB/s").

Times the fused pool/header/match kernel (memvul_pool_match, MEMVUL_PM_ALL) and the match phase alone on
  * the bank-streaming regime (few queries per pass over a bank larger than L2): HBM-bound, the >=60 % target;
  * BASELINE config 4 (B=256, G=16,384) and config 2 (B=64, G=129): FP32-ALU / latency bound.
Algorithmic bytes (fp32): 4*(G*512 + B*512 + 2*B*G) + 4*2*B*G for the probs tensor the kernel also writes.
Prints one JSON object per case; bench.py embeds the streaming case in its line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N  # noqa: E402

H, D = 768, 512


def run_case(B, G, iters=20, phase=None):
    dev = "cuda"
    torch.manual_seed(0)
    cls = torch.randn(B, H, device=dev)
    wp, bp = torch.randn(H, H, device=dev) * 0.03, torch.randn(H, device=dev) * 0.02
    wh, bh = torch.randn(D, H, device=dev) * 0.03, torch.randn(D, device=dev) * 0.02
    wproj = torch.randn(2, 3 * D, device=dev) * 0.03
    bank = torch.relu(torch.randn(G, D, device=dev))
    vterm = N.bank_prepare(bank, wproj)
    u = torch.relu(torch.randn(B, D, device=dev))
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)       # 256 MB > 126 MB L2
    flush_sink = torch.zeros(1, dtype=torch.float32, device=dev)
    mask = N.PM_ALL if phase is None else phase

    def call():
        return N.pool_match(cls, H, B, wp, bp, wh, bh, wproj, bank, vterm, phase_mask=mask, u=u)
    for _ in range(3):
        call()
    ms = []
    for _ in range(iters):
        flush_sink.copy_(flush[:1] + flush.sum())      # evict the bank from L2 with CLEAN lines (a memset would leave
                                                        # 126 MB of dirty lines whose write-back competes with the stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    t = ms[len(ms) // 2]
    bytes_alg = 4 * (G * D + B * D + 2 * B * G) + 4 * 2 * B * G
    lane_instr = 3.0 * B * G * D
    return {"B": B, "G": G, "phases": "all" if phase is None else "match+final", "us": t * 1e3,
            "algorithmic_MB": bytes_alg / 1e6, "hbm_GBps": bytes_alg / (t * 1e-3) / 1e9,
            "fp32_lane_Tinstr_per_s": lane_instr / (t * 1e-3) / 1e12, "l2": "flushed by reading a 256 MB buffer between launches"}


if __name__ == "__main__":
    cases = [(1, 262144), (2, 262144), (4, 262144), (4, 65536), (8, 65536), (256, 16384), (64, 129)]
    if "--c4" in sys.argv:
        cases = [(256, 16384), (1024, 16384)]
    for B, G in cases:
        for ph in (None, N.PM_UTERM | N.PM_MATCH | N.PM_FINAL):
            d = run_case(B, G, phase=ph)
            if "--table" in sys.argv:
                print(f"B={d['B']:4d} G={d['G']:7d} {d['phases']:12s} {d['us']:9.1f} us {d['hbm_GBps']:8.1f} GB/s "
                      f"{d['fp32_lane_Tinstr_per_s']:.2f} Tlane-instr/s", flush=True)
            else:
                print(json.dumps(d), flush=True)
