"""Parity gates of SURVEY.md 8d ("parity gates per run") as one report: given this library's outputs and the outputs of
a checker for the same inputs (the tests and ``bench.py`` pass the CPU restatement of the reference; this module never
computes a reference itself), measure

  * ``max_logit_err``            max |logits - ref|                       gate: <= 1e-3 (BASELINE.json north_star)
  * labels at each threshold      ``pos`` iff max_g P(same) >= thres       (predict_memory.py:168-177), compared on ALL rows
  * ``min_margin``                min_b |max_g P_ref(same) - thres|: how far the closest row is from the decision boundary
  * ``rows_excluded``             rows whose reference margin is <= tol at some threshold, i.e. rows on which a label
                                  difference would be within the stated logit tolerance; 0 means every label was gated
  * arg-max anchor                identical on every row whose reference top-2 gap exceeds 2*tol (``argmax_clear_rows``),
                                  and on ALL rows the chosen anchor must be a maximiser up to the observed probability error
"""
from __future__ import annotations

from typing import Dict, Iterable

import numpy as np


def gate_report(logits, probs, best_idx, ref_logits, ref_probs, same_idx: int,
                thresholds: Iterable[float] = (0.5,), tol: float = 1e-3) -> Dict[str, object]:
    logits, probs = np.asarray(logits, dtype=np.float64), np.asarray(probs, dtype=np.float64)
    ref_logits, ref_probs = np.asarray(ref_logits, dtype=np.float64), np.asarray(ref_probs, dtype=np.float64)
    best_idx = np.asarray(best_idx).astype(np.int64)
    B, G = probs.shape[0], probs.shape[1]
    ps, ps_ref = probs[:, :, same_idx], ref_probs[:, :, same_idx]
    vote, vote_ref = ps.max(1), ps_ref.max(1)
    rep: Dict[str, object] = {"rows": int(B), "anchors": int(G), "tol": tol,
                              "max_logit_err": float(np.abs(logits - ref_logits).max()),
                              "max_prob_err": float(np.abs(probs - ref_probs).max())}
    excluded = np.zeros(B, dtype=bool)
    labels = {}
    for t in thresholds:
        margin = np.abs(vote_ref - t)
        within = margin <= tol
        excluded |= within
        mism = (vote >= t) != (vote_ref >= t)
        labels[f"{t:g}"] = {"mismatch_rows": int(mism.sum()), "mismatch_outside_tol": int((mism & ~within).sum()),
                            "min_margin": float(margin.min()), "pos_ref": int((vote_ref >= t).sum())}
    rep["labels"] = labels
    rep["min_margin"] = min(v["min_margin"] for v in labels.values())
    rep["rows_excluded"] = int(excluded.sum())
    rep["label_mismatch_outside_tol"] = int(sum(v["mismatch_outside_tol"] for v in labels.values()))
    ref_idx = ps_ref.argmax(1)
    if G > 1:
        srt = np.sort(ps_ref, axis=1)
        gap = srt[:, -1] - srt[:, -2]
    else:
        gap = np.ones(B)
    clear = gap > 2 * tol
    rep["argmax_clear_rows"] = int(clear.sum())
    rep["argmax_mismatch_clear"] = int((best_idx[clear] != ref_idx[clear]).sum())
    rep["argmax_mismatch_all"] = int((best_idx != ref_idx).sum())
    # every row, near-ties included: the chosen anchor is a maximiser of the REFERENCE probabilities up to twice the
    # observed probability error
    chosen = ps_ref[np.arange(B), best_idx]
    rep["argmax_not_maximiser"] = int((chosen < vote_ref - 2 * max(rep["max_prob_err"], 1e-7)).sum())
    rep["ok"] = bool(rep["max_logit_err"] <= tol and rep["label_mismatch_outside_tol"] == 0
                     and rep["argmax_mismatch_clear"] == 0 and rep["argmax_not_maximiser"] == 0)
    return rep
