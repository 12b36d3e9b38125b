"""Writes tests/golden/*.npz: seeded inputs and the CPU oracle's outputs for them.

The reference ships no golden vectors and cannot be imported here, so these are
self-generated known answers: they pin the oracle against drift (tests/test_oracle.py recomputes them on
CPU) and are what the GPU parity tests compare the CUDA path with.  Run:  python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import memvul_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (shape, issue lens, padded S, anchor lens, same_idx)
    "tiny_ragged": (O.BERT_TINY, [48, 17, 2, 33, 48, 9], 48, [30, 12, 48, 5, 21, 40, 7, 16, 25], 0),
    "tiny_same1": (O.BERT_TINY, [130, 200, 64], 200, [64, 150, 9, 77], 1),
    "base_small": (O.BERT_BASE, [64, 23, 40, 57], 64, [40, 18, 64, 9, 33, 50, 27, 12, 45, 60, 21, 36], 0),
}


def make(name):
    shape, lens, S, alens, same = CASES[name]
    torch.manual_seed(0)
    sd = O.synthetic_state_dict(shape)
    ids, mask, tids = O.synthetic_ids(len(lens), S, lens=lens, seed=11, vocab_size=shape.vocab_size)
    a_ids, a_mask, _ = O.synthetic_ids(len(alens), max(alens), lens=alens, seed=12, vocab_size=shape.vocab_size)
    anchors = [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(len(alens))]
    with torch.no_grad():
        bank = O.build_bank(sd, anchors, shape, chunk=128)
        out = O.memory_forward(sd, ids, mask, tids, bank, same, shape)
    wsum = float(sum(v.double().sum() for v in sd.values()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ids=ids.numpy(), mask=mask.numpy(), anchor_ids=a_ids.numpy(),
                        anchor_mask=a_mask.numpy(), same_idx=same, bank=bank.numpy(), u=out["u"].numpy(),
                        logits=out["logits"].numpy(), p=out["p"].numpy(), best_idx=out["best_idx"].numpy(),
                        probs=out["probs"].numpy(), weight_checksum=wsum)
    print(name, "logits", tuple(out["logits"].shape), "best", out["best_idx"].tolist(), "checksum", wsum)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for n in CASES:
        make(n)
