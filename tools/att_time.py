import os, sys, torch
sys.path.insert(0, "/root/repo")
from memvul_b200 import native as N
H=768
for (B,S) in [(64,512),(256,128),(128,256),(512,64)]:
    qkv = torch.randn(B*S, 3*H, device="cuda").half()
    lens = torch.full((B,), S, dtype=torch.int32, device="cuda")
    for _ in range(3): N.attention_f16(qkv, lens, B, S, H)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): N.attention_f16(qkv, lens, B, S, H)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    nct = B*12*((S+127)//128); nkb=(S+63)//64
    print(f"B={B} S={S}: {ms*1e3:.1f} us; CTAs {nct}, key blocks/CTA {nkb}; per CTA-slot {ms*1e3/(nct/296):.2f} us")
