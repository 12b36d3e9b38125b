"""Debug: phase timeline of ONE attention CTA (CTA 0) from clock64() stamps (MEMVUL_ATT_TRACE=<file>).
    MEMVUL_ATT_TRACE=/tmp/att.bin python tools/att_trace.py
Soft-max warp 0 slots per key block g: 0 loop top, 1 S landed, 2 S in registers, 3 row max done, 4 exp/sum/pack/STS done,
5 pv_done(g-1) waited (+ rescale), 6 P handed to the MMA warp, 7 (last block of an item) final P.V retired.
MMA warp slots: 0 P_g landed, 1 V_g landed, 2 P.V issued, 3 Q.K^T(g+2) issued."""
import os, struct, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N
path = os.environ["MEMVUL_ATT_TRACE"]
B, S, H = 64, 512, 768
qkv = torch.randn(B * S, 3 * H, device="cuda").half()
lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
for _ in range(3): N.attention_f16(qkv, lens, B, S, H)
torch.cuda.synchronize()
raw = open(path, "rb").read()
v = struct.unpack("<2048Q", raw)
sm = [[v[g * 8 + k] for k in range(8)] for g in range(128)]
mm = [[v[1024 + g * 8 + k] for k in range(8)] for g in range(128)]
t0 = sm[0][0]
names = ["wait S", "ld S", "row max", "exp+pack+STS", "wait pv(g-1)", "fence+arrive"]
print("soft-max warp 0 of CTA 0, cycles per phase (key blocks 8..71 = items 1..8):")
acc = [0.0] * 6; period = 0.0; n = 0; item_gap = []
for g in range(8, 72):
    if not sm[g][6] or not sm[g + 1][0]: continue
    for k in range(6): acc[k] += sm[g][k + 1] - sm[g][k]
    period += sm[g + 1][0] - sm[g][0]; n += 1
    if g % 8 == 7: item_gap.append(sm[g + 1][0] - sm[g][6])
for k in range(6): print(f"  {names[k]:>14}: {acc[k]/n:7.0f}")
print(f"  {'block period':>14}: {period/n:7.0f}   (loop top to loop top, includes the item epilogue/prologue every 8th block: {sum(item_gap)/max(1,len(item_gap)):.0f} cycles)")
lag = [mm[g][0] - sm[g][6] for g in range(8, 72) if mm[g][0] and sm[g][6]]
pv = [mm[g][2] - mm[g][0] for g in range(8, 72) if mm[g][2]]
qk = [mm[g][3] - mm[g][2] for g in range(8, 72) if mm[g][3]]
print(f"MMA warp: P handed over -> seen {sum(lag)/len(lag):.0f} (warp 0's arrive; the barrier needs all 4 warps), P.V issue {sum(pv)/len(pv):.0f}, Q.K^T(g+2) issue {sum(qk)/len(qk):.0f}")
print("first blocks (cycles since start): " + ", ".join(f"g{g}:{sm[g][0]-t0}" for g in range(0, 17)))
print("item transitions (cycles): last arrive -> final P.V retired -> next loop top (O read-out + decode) -> first S landed -> first block done")
for g in range(7, 64, 8):
    a, b, c, d, e = sm[g][6], sm[g][7], sm[g + 1][0], sm[g + 1][1], sm[g + 1][6]
    m = mm[g]
    print(f"  g={g}: {b-a:5d} {c-b:5d} {d-c:5d} {e-d:5d} | MMA: P_g seen {m[0]-a:5d} after the arrive, P.V issued +{m[2]-m[0]}, next item's P_0 seen at +{mm[g+1][0]-a}")
print("O read-out detail (cycles): final P.V retired -> O in registers -> o_free arrived -> ctx stored -> next item decoded (lens[b] landed) -> loop top")
for g in range(7, 64, 8):
    b, t4, t5, t6, t7, c = sm[g][7], mm[g][4], mm[g][5], mm[g][6], mm[g + 1][7], sm[g + 1][0]
    print(f"  g={g}: {t4-b:5d} {t5-t4:5d} {t6-t5:5d} {t7-t6:5d} {c-t7:5d}")
