"""BASELINE configs[3] end to end: 16,384 synthetic anchors (bank = relu(N(0,1))), bert-base S=512, batch 256."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200 import native
from memvul_b200.synthetic import BERT_BASE, build_memory_model, synthetic_ids
dev = torch.device("cuda:0")
model, _ = build_memory_model(BERT_BASE, device=dev)
G, B, S = 16384, 256, 512
g = torch.Generator().manual_seed(4)
model._golden_instances_embeddings = torch.relu(torch.randn(G, 512, generator=g)).to(dev)
model._golden_instances_labels = [f"CWE-{i}" for i in range(G)]
ids, mask, tids = synthetic_ids(B, S, seed=44)
sample = {"tokens": {"token_ids": ids.to(dev), "mask": mask.to(dev), "type_ids": tids.to(dev)}}
with torch.no_grad():
    for _ in range(2):
        r = model.match_batch(sample)
    native.profile_enable(True); native.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = model.match_batch(sample)
    e1.record(); torch.cuda.synchronize()
    prof = native.profile_read(); native.profile_enable(False)
ms = e0.elapsed_time(e1) / 3
p = r["probs"]
assert torch.equal(r["best_idx"].long(), p[:, :, model._same_idx].argmax(1))
print(f"config4: B={B} S={S} G={G}: {ms:.2f} ms/batch -> {B/ms*1e3:.0f} issues/s; pool_match {prof['pool_match']['ms']/3:.3f} ms; "
      f"probs {tuple(p.shape)} finite={bool(torch.isfinite(p).all())}; mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
