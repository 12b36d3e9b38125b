"""Throw-away numerics study (CPU): how far do the match logits move when the encoder GEMM
operands are rounded to fp16 / bf16 / tf32 (fp32 accumulate), with an fp32 or fp16 residual
stream?  Decides the operand format for the tcgen05 kernels (DESIGN.md "Precision").
Uses the oracle as the fp32 truth; it is a study tool, not product code."""
import math, sys, torch
sys.path.insert(0, ".")
from oracle import memvul_oracle as O

def rnd(x, fmt):
    if fmt == "fp32": return x
    if fmt == "fp16": return x.half().float()
    if fmt == "bf16": return x.bfloat16().float()
    if fmt == "tf32":
        i = x.view(torch.int32); i = (i + 0x1000) & ~0x1FFF  # round to 10-bit mantissa
        return i.view(torch.float32)
    raise ValueError(fmt)

def enc(sd, ids, maskf, shape, fmt, resid16=False, p16=True):
    B, S = ids.shape; H, nH, dh = shape.hidden, shape.heads, shape.head_dim
    e = O.EMB + "embeddings."
    x = sd[e+"word_embeddings.weight"][ids] + sd[e+"position_embeddings.weight"][:S][None] + sd[e+"token_type_embeddings.weight"][0]
    x = O._ln(x, sd[e+"LayerNorm.weight"], sd[e+"LayerNorm.bias"], shape.ln_eps)
    ext = (1.0 - maskf)[:, None, None, :] * -10000.0
    lin = lambda a, w, b: torch.nn.functional.linear(rnd(a, fmt), rnd(w, fmt), b)
    for l in range(shape.layers):
        p = O.EMB + f"encoder.layer.{l}."
        if resid16: x = rnd(x, fmt)
        q = rnd(lin(x, sd[p+"attention.self.query.weight"], sd[p+"attention.self.query.bias"]), fmt)
        k = rnd(lin(x, sd[p+"attention.self.key.weight"], sd[p+"attention.self.key.bias"]), fmt)
        v = rnd(lin(x, sd[p+"attention.self.value.weight"], sd[p+"attention.self.value.bias"]), fmt)
        q = q.view(B,S,nH,dh).transpose(1,2); k = k.view(B,S,nH,dh).transpose(1,2); v = v.view(B,S,nH,dh).transpose(1,2)
        sc = q @ k.transpose(-1,-2) / math.sqrt(dh) + ext
        m = sc.max(-1, keepdim=True).values
        pe = torch.exp(sc - m); den = pe.sum(-1, keepdim=True)
        if p16: pe = rnd(pe, fmt)
        ctx = rnd(((pe @ v) / den).transpose(1,2).reshape(B,S,H), fmt)
        a = lin(ctx, sd[p+"attention.output.dense.weight"], sd[p+"attention.output.dense.bias"])
        x = O._ln(a + x, sd[p+"attention.output.LayerNorm.weight"], sd[p+"attention.output.LayerNorm.bias"], shape.ln_eps)
        if resid16: x = rnd(x, fmt)
        h = rnd(O._gelu_erf(lin(x, sd[p+"intermediate.dense.weight"], sd[p+"intermediate.dense.bias"])), fmt)
        o = lin(h, sd[p+"output.dense.weight"], sd[p+"output.dense.bias"])
        x = O._ln(o + x, sd[p+"output.LayerNorm.weight"], sd[p+"output.LayerNorm.bias"], shape.ln_eps)
    return x

if __name__ == "__main__":
    torch.manual_seed(0)
    shape = O.BERT_BASE
    sd = O.synthetic_state_dict(shape)
    B, S = 6, 128
    ids, mask, _ = O.synthetic_ids(B, S, lens=[128, 100, 64, 128, 17, 90])
    bank = torch.relu(torch.randn(129, 512) * 0.5)
    lin = torch.nn.functional.linear
    def head(hid):
        pooled = torch.tanh(lin(hid[:,0], sd["_bert_pooler.pooler.dense.weight"], sd["_bert_pooler.pooler.dense.bias"]))
        u = torch.relu(lin(pooled, sd["_projector_single._linear_layers.0.weight"], sd["_projector_single._linear_layers.0.bias"]))
        return u, O.match(u, bank, sd["_projector.weight"], 0)
    ref = enc(sd, ids, mask.float(), shape, "fp32")
    u0, m0 = head(ref)
    print("logit range", m0["logits"].min().item(), m0["logits"].max().item(), "u max", u0.max().item())
    for fmt, r16 in [("fp16", False), ("fp16", True), ("tf32", False), ("bf16", False)]:
        h = enc(sd, ids, mask.float(), shape, fmt, resid16=r16)
        u, m = head(h)
        print(f"{fmt:5s} resid16={r16}: cls max|d|={(h[:,0]-ref[:,0]).abs().max():.3e}  u max|d|={(u-u0).abs().max():.3e}  "
              f"logits max|d|={(m['logits']-m0['logits']).abs().max():.3e}  idx same={bool((m['best_idx']==m0['best_idx']).all())}")

def study_default_init():
    """Same study with the SURVEY 8c init: N(0,0.02) everywhere, LN gamma=1 beta=0, zero biases,
    header/projector PyTorch default Linear init."""
    shape = O.BERT_BASE
    sd = O.synthetic_state_dict(shape)
    g = torch.Generator().manual_seed(7)
    for k in sd:
        if "LayerNorm.weight" in k: sd[k] = torch.ones_like(sd[k])
        elif "LayerNorm.bias" in k or k.endswith(".bias"): sd[k] = torch.zeros_like(sd[k])
        elif k.startswith(O.EMB): sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
    for k in ["_bert_pooler.pooler.dense.weight", "_projector_single._linear_layers.0.weight", "_projector.weight"]:
        fan_in = sd[k].shape[1]; b = 1 / math.sqrt(fan_in)
        sd[k] = (torch.rand(sd[k].shape, generator=g) * 2 - 1) * b
    return sd
