"""Offline helper (CPU, build container): evaluates the ORACLE on the inputs a parity test will use and prints the
decision margins, so test inputs can be chosen with guaranteed margins (no row within 2*TOL of a threshold) and the
expected number of clear arg-max rows can be pinned in the test.  Not imported by the product or the tests."""
import argparse
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import memvul_oracle as O  # noqa: E402


sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from config_inputs import c2_inputs, split_threshold  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs="+", default=[9])
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--head-scale", type=float, default=1.0)
    args = ap.parse_args()
    sd = O.synthetic_state_dict(O.BERT_BASE, 2021)
    if args.head_scale != 1.0:
        sd["_projector.weight"] = sd["_projector.weight"] * args.head_scale
    for seed in args.seeds:
        t0 = time.time()
        a_ids, a_mask, alens, ids, mask, tids, lens = c2_inputs(seed, args.B)
        with torch.no_grad():
            bank = O.build_bank(sd, [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(len(alens))])
            ref = O.memory_forward(sd, ids, mask, tids, bank, 0)
        ps = ref["p"][:, :, 0]
        vote = ps.max(1).values
        top2 = ps.topk(2, dim=1).values
        gap = top2[:, 0] - top2[:, 1]
        print(f"seed {seed}: {time.time() - t0:.0f}s vote min/max {float(vote.min()):.4f}/{float(vote.max()):.4f} "
              f"min|vote-0.5| {float((vote - 0.5).abs().min()):.2e} pos@0.5 {int((vote >= 0.5).sum())} "
              f"top2 gap: min {float(gap.min()):.2e} median {float(gap.median()):.2e} clear(>2e-3) {int((gap > 2e-3).sum())}/{len(gap)} "
              f"logit absmax {float(ref['logits'].abs().max()):.3f} split threshold/margin {split_threshold(vote)}", flush=True)


if __name__ == "__main__":
    main()
