"""One-process correctness + timing check of the attention kernel selected by MEMVUL_ATT_V / MEMVUL_ATT_POLY
(tools/gpu_att3.sh runs it once per variant).  Padded and packed (row_start) layouts against an fp32 torch
reference on the same fp16-rounded inputs, then CUDA-event timing of the bert-base shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native as N  # noqa: E402


def ref_rows(qkv, lens, B, S, H, row_start=None):
    nh = H // 64
    q, k, v = qkv.float().split(H, dim=1)
    out = torch.zeros(qkv.shape[0], H, device=qkv.device)
    valid = torch.zeros(qkv.shape[0], dtype=torch.bool, device=qkv.device)
    for b in range(B):
        r0 = int(row_start[b]) if row_start is not None else b * S
        L = int(lens[b])
        valid[r0:r0 + L] = True
        qq = q[r0:r0 + L].view(L, nh, 64).transpose(0, 1)
        kk = k[r0:r0 + L].view(L, nh, 64).transpose(0, 1)
        vv = v[r0:r0 + L].view(L, nh, 64).transpose(0, 1)
        s = qq @ kk.transpose(1, 2) / 8.0
        out[r0:r0 + L] = (torch.softmax(s, -1) @ vv).transpose(0, 1).reshape(L, H)
    return out, valid


def main():
    tag = f"ATT_V={os.environ.get('MEMVUL_ATT_V', '1')} POLY={os.environ.get('MEMVUL_ATT_POLY', '0')}"
    ok = True
    torch.manual_seed(1)
    cases = [(1, 128, 128, [128], False), (2, 128, 128, [128, 77], False), (2, 256, 128, [256, 130], False),
             (3, 512, 768, [512, 300, 5], False), (2, 200, 768, [200, 129], False), (4, 64, 128, [64, 2, 33, 17], False),
             (8, 512, 768, [512] * 8, False), (1, 1, 128, [1], False), (2, 512, 128, [511, 512], False),
             (3, 384, 128, [384, 200, 129], False), (4, 384, 768, [384, 257, 256, 1], True), (2, 320, 128, [320, 127], False),
             (5, 512, 128, [512, 1, 130, 64, 300], True), (6, 511, 128, [257, 511, 129, 63, 200, 31], True),
             (40, 512, 768, [512, 77, 300, 128, 129] * 8, True), (64, 512, 768, [512] * 64, False)]
    for (B, S, H, lens, packed) in cases:
        scale = 1.5 if B < 40 else 1.0
        qkv = (torch.randn(B * S, 3 * H, device="cuda") * scale).half()
        lens_t = torch.tensor(lens, dtype=torch.int32, device="cuda")
        rs = None
        if packed:
            rs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
        ctx = N.attention_f16(qkv, lens_t, B, S, H, row_start=rs)
        torch.cuda.synchronize()
        ref, valid = ref_rows(qkv, lens, B, S, H, rs.cpu() if packed else None)
        d = (ctx.float() - ref)[valid].abs()
        bad = bool(torch.isnan(ctx.float()[valid]).any()) or float(d.max()) > 4e-3
        if packed and int(valid.sum()) < B * S:
            bad |= float(ctx[int(valid.sum()):].float().abs().max()) != 0.0
        ok &= not bad
        print(f"  {'FAIL' if bad else 'ok  '} [{tag}] B={B} S={S} H={H} packed={packed}: max|d|={float(d.max()):.2e} "
              f"mean|d|={float(d.mean()):.2e}", flush=True)
    H = 768
    for (B, S) in [(64, 512), (256, 128), (128, 256)]:
        qkv = torch.randn(B * S, 3 * H, device="cuda").half()
        lens_t = torch.full((B,), S, dtype=torch.int32, device="cuda")
        for _ in range(5):
            N.attention_f16(qkv, lens_t, B, S, H)
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                N.attention_f16(qkv, lens_t, B, S, H)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        # N.attention_f16 also zero-fills ctx (a 50 MB memset per call): identical for every variant
        print(f"  time [{tag}] B={B} S={S}: {best * 1e3:.1f} us (incl. the wrapper's ctx memset)", flush=True)
    print(f"RESULT [{tag}] {'PASS' if ok else 'FAIL'}", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
