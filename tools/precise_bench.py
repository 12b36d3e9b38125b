"""Throughput and per-kernel-class time of the opt-in accuracy mode (precision="split_fp16", MEMVUL_ENC_PRECISE) next to
the default fp16-operand path, same C2-shaped batch (bert-base, S=512, B issue reports, 129 anchors), plus the logit
difference between the two on that batch.

    python tools/precise_bench.py [--batch 64] [--steps 5] [--out profiles/r02n_precise_bench.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memvul_b200 import native as N  # noqa: E402
from memvul_b200.synthetic import BERT_BASE, build_memory_model, synthetic_ids  # noqa: E402


def run(precision, B, steps, dev):
    model, _ = build_memory_model(BERT_BASE, device=dev, precision=precision)
    g = torch.Generator().manual_seed(2022)
    a_lens = torch.randint(64, 513, (129,), generator=g).tolist()
    with torch.no_grad():
        for c0, c1 in ((0, 128), (128, 129)):
            lens_c = a_lens[c0:c1]
            ids, mask, tids = synthetic_ids(len(lens_c), max(lens_c), lens=lens_c, seed=2023 + c0)
            model.forward_gold_instances({"tokens": {"token_ids": ids.to(dev), "mask": mask.to(dev), "type_ids": tids.to(dev)}},
                                         [{"type": "golden", "instance": [{"label": f"CWE-{c0 + i}"}]} for i in range(len(lens_c))])
        ids, mask, tids = synthetic_ids(B, 512, seed=2121)
        sample = {"tokens": {"token_ids": ids.to(dev), "mask": mask.to(dev), "type_ids": tids.to(dev)}}
        for _ in range(2):
            out = model.match_batch(sample)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = model.match_batch(sample)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        N.profile_enable(True)
        N.profile_read()
        for _ in range(2):
            model.match_batch(sample)
        prof = N.profile_read()
        N.profile_enable(False)
    classes = {k: {"ms_per_step": round(v["ms"] / 2, 3), "launches_per_step": v["launches"] / 2} for k, v in prof.items() if v["launches"]}
    res = {"precision": precision, "ms_per_step": round(ms, 3), "issues_per_s": round(B / ms * 1e3, 1), "classes": classes}
    return res, out["logits"].double().cpu()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    fast, lg_fast = run("fp16", args.batch, max(args.steps, 20), dev)
    torch.cuda.empty_cache()
    prec, lg_prec = run("split_fp16", args.batch, args.steps, dev)
    rep = {"workload": f"bert-base, S=512, {args.batch} issue reports, 129 anchors (C2 shape), resident inputs",
           "fp16": fast, "split_fp16": prec, "slowdown": round(prec["ms_per_step"] / fast["ms_per_step"], 2),
           "max_abs_logit_diff_between_modes": float((lg_fast - lg_prec).abs().max())}
    print(json.dumps(rep, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
