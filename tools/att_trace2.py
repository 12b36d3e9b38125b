"""Debug: phase timeline of ONE CTA of attention_tcgen05_v2 (MEMVUL_ATT_TRACE=<file>): soft-max warp 0 and the P.V warp.
Soft-max slots per key block g: 0 loop top, 1 S landed, 2 row max done, 3 exp/sum/pack/STS done, 4 P handed over;
last block of an item: 5 final P.V retired, 6 item epilogue done.  P.V warp: 0 P_g seen, 2 P.V issued."""
import os, struct, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N
path = os.environ["MEMVUL_ATT_TRACE"]
B, S, H = 64, 512, 768
qkv = torch.randn(B * S, 3 * H, device="cuda").half()
lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
for _ in range(3): N.attention_f16(qkv, lens, B, S, H)
torch.cuda.synchronize()
v = struct.unpack("<2048Q", open(path, "rb").read())
sm = [[v[g * 8 + k] for k in range(8)] for g in range(128)]
mm = [[v[1024 + g * 8 + k] for k in range(8)] for g in range(128)]
names = ["wait P free + S", "ld + row max", "exp+pack+STS", "rescale+fence+arrive"]
acc = [0.0] * 4; n = 0; period = 0.0
for g in range(8, 71):
    if not sm[g][4] or not sm[g + 1][0]: continue
    for k in range(4): acc[k] += sm[g][k + 1] - sm[g][k]
    period += sm[g + 1][0] - sm[g][0]; n += 1
print("soft-max warp 0 of CTA 0, mean cycles per phase over key blocks 8..70:")
for k in range(4): print(f"  {names[k]:>22}: {acc[k] / max(n, 1):7.0f}")
print(f"  {'block period':>22}: {period / max(n, 1):7.0f}")
lag = [mm[g][0] - sm[g][4] for g in range(8, 72) if mm[g][0] and sm[g][4]]
pv = [mm[g][2] - mm[g][0] for g in range(8, 72) if mm[g][2]]
print(f"P.V warp: P handed over (warp 0) -> seen {sum(lag) / max(len(lag), 1):.0f}, P.V issue {sum(pv) / max(len(pv), 1):.0f}")
print("item boundaries (cycles): last hand-over -> final P.V retired -> epilogue done -> next item's first S landed")
for g in range(7, 64, 8):
    print(f"  g={g}: {sm[g][5] - sm[g][4]:6d} {sm[g][6] - sm[g][5]:6d} {sm[g + 1][1] - sm[g][6]:6d}   item period {sm[g + 8][0] - sm[g][0] if sm[g + 8][0] else 0}")
