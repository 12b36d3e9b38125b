"""Debug: phase timeline of the fused GEMM+LayerNorm kernel's CTA 0 (epilogue warp 0 + MMA warp) from clock64() stamps.
    MEMVUL_LN_TRACE=/tmp/ln.bin python tools/ln_trace.py [K ...]
Epilogue slots per tile: 0 loop top, 1 accumulator ready, 2-5 residual chunk c landed, 6 pass 1 done, 7 stats published
(+ TMEM stores retired), 8 stats of all column blocks arrived, 9-12 pass-2 chunk c handed to the TMA store, 13 tile done
(accumulator released, staging drained, next residuals requested).  MMA warp: 14 tile start, 15 last MMA issued."""
import os, struct, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N
path = os.environ["MEMVUL_LN_TRACE"]
M = 32768
for K in [int(x) for x in sys.argv[1:]] or [768, 3072]:
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(768, K, device="cuda") * 0.05).half()
    bias = torch.randn(768, device="cuda"); resid = torch.randn(M, 768, device="cuda")
    g = torch.rand(768, device="cuda") + 0.5; b = torch.randn(768, device="cuda")
    for _ in range(3): N.gemm_ln_f16(a, w, bias, resid, g, b, inplace=True)
    torch.cuda.synchronize()
    v = struct.unpack("<128Q", open(path, "rb").read())
    t = [[v[i * 16 + k] for k in range(16)] for i in range(8)]
    t0 = min(x for row in t for x in row if x)
    print(f"K={K}: cycles since the first stamp (tile: slots 0..13 | MMA start, MMA issued)")
    for i, row in enumerate(t):
        if not row[0]: continue
        print(f"  tile {i}: " + " ".join(f"{x - t0:7d}" if x else "      -" for x in row[:14]) + " | " + " ".join(f"{x - t0:7d}" if x else "      -" for x in row[14:]))
    names = ["wait acc", "res0", "res1", "res2", "res3", "pass1 tail", "publish+st", "wait stats", "p2 c0", "p2 c1", "p2 c2", "p2 c3", "release+drain"]
    n = 0; acc = [0] * 13
    for row in t[1:]:
        if not row[13]: continue
        n += 1
        for k in range(13): acc[k] += row[k + 1] - row[k]
    if n:
        print("  mean cycles per phase (tiles 1..): " + ", ".join(f"{names[k]} {acc[k] // n}" for k in range(13)) + f" | tile period {sum(acc) // n}")
