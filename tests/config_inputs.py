"""Seeded inputs of the BASELINE-config parity tests, shared with tools/margin_probe.py (which evaluates the CPU oracle
on them offline so that the tests can pin decision margins instead of gating on possibly empty masks)."""
import torch

from memvul_b200.synthetic import synthetic_ids

C2_SEED = 12       # tools/margin_probe.py: 63 of 64 rows have an arg-max gap > 2e-3, min |vote - 0.5| = 8.7e-3


def c2_inputs(seed: int = C2_SEED, B: int = 64, G: int = 129):
    """BASELINE configs[1] shape: 64 issue reports padded to 512 (every other one full length, the rest 300..512) and
    129 anchors of 16..128 tokens."""
    g = torch.Generator().manual_seed(seed)
    alens = torch.randint(16, 129, (G,), generator=g).tolist()
    a_ids, a_mask, _ = synthetic_ids(G, 128, lens=alens, seed=seed + 1)
    lens = torch.randint(300, 513, (B,), generator=g).tolist()
    for i in range(0, B, 2):
        lens[i] = 512
    ids, mask, tids = synthetic_ids(B, 512, lens=lens, seed=seed + 2)
    return a_ids, a_mask, alens, ids, mask, tids, lens


def c5_inputs(seed: int = 5):
    """Mixed-length stream in data order: lengths from {128, 256, 512} plus ragged neighbours, 16 short anchors."""
    lens = [512, 128, 256, 100, 512, 256, 128, 500, 200, 128, 256, 512]
    ids, mask, tids = synthetic_ids(len(lens), 512, lens=lens, seed=seed)
    alens = [12, 40, 64, 9, 33, 64, 20, 50, 64, 17, 8, 64, 30, 45, 25, 60]
    a_ids, a_mask, _ = synthetic_ids(len(alens), 64, lens=alens, seed=seed + 1)
    return a_ids, a_mask, alens, ids, mask, tids, lens


def split_threshold(votes: torch.Tensor):
    """A threshold that separates the rows into two classes with the largest possible margin: the midpoint of the widest
    gap between consecutive sorted votes in the middle half.  Returns (threshold, margin)."""
    v = torch.sort(votes.double()).values
    n = v.numel()
    lo, hi = n // 4, max(n // 4 + 1, 3 * n // 4)
    gaps = v[lo + 1:hi + 1] - v[lo:hi]
    k = int(torch.argmax(gaps))
    return float((v[lo + k] + v[lo + k + 1]) / 2), float(gaps[k] / 2)
