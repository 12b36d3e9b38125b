"""Host-side metrics fed by the match kernel's per-sample result.

``SiameseMeasureV1`` mirrors MemVul/custom_metric.py:55-98 (registered ``siamese_measure_v1``):
it records (is-positive, P(same) at the arg-max anchor) per sample and, on ``get_metric(reset=True)``,
sweeps thresholds 0.50..0.89 (the reference's ``find_best_thres``, :35-52 -- the LAST threshold that
reaches the best F1 wins because of its ``>=``) and reports ROC-AUC / average precision (:88-90).
The sweep is vectorised with numpy; results are value-identical to the reference's python loops.
``CategoricalAccuracy`` / ``FBetaMeasure`` are the small subset of the AllenNLP metrics that
``ModelMemory.get_metrics`` reads (model_memory.py:80-85,194-203), operating on host arrays.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .registrable import Metric


def confusion(labels: np.ndarray, pred: np.ndarray) -> Dict[str, float]:
    labels = np.asarray(labels).astype(bool)
    pred = np.asarray(pred).astype(bool)
    tp = int(np.sum(pred & labels)); fn = int(np.sum(~pred & labels))
    tn = int(np.sum(~pred & ~labels)); fp = int(np.sum(pred & ~labels))
    recall = tp / (tp + fn) if tp + fn else 0
    prec = tp / (tp + fp) if tp + fp else 0
    f1 = 2 * recall * prec / (recall + prec) if recall + prec else 0
    return {"TP": tp, "FN": fn, "TN": tn, "FP": fp, "precision": prec, "recall": recall, "f1": f1}


def find_best_thres(labels: Sequence[int], scores: Sequence[float], interval=(0.5, 0.9)) -> Dict[str, float]:
    labels = np.asarray(labels)
    scores = np.asarray(scores, dtype=np.float64)
    best, best_f1 = None, 0
    for thres in np.arange(interval[0], interval[1], 0.01):
        m = confusion(labels, scores >= thres)
        if m["f1"] >= best_f1:
            best_f1 = m["f1"]
            m["thres"] = thres
            best = m
    return best


def roc_auc(labels: Sequence[int], scores: Sequence[float]) -> float:
    from sklearn import metrics
    fpr, tpr, _ = metrics.roc_curve(labels, scores, pos_label=1)
    return float(metrics.auc(fpr, tpr))


def average_precision(labels: Sequence[int], scores: Sequence[float]) -> float:
    from sklearn import metrics
    return float(metrics.average_precision_score(labels, scores, pos_label=1))


@Metric.register("siamese_measure_v1")
class SiameseMeasureV1(Metric):
    def __init__(self, same_idx: int, thres: float = 0.5) -> None:
        self._same_idx = same_idx
        self._thres = thres
        self._labels: List[int] = []
        self._scores: List[float] = []

    def __call__(self, predictions, metadata: List[Dict[str, Any]] = None, mask=None) -> None:
        """predictions: [B,2] probabilities at the arg-max anchor (tensor, ndarray or nested list)."""
        probs = predictions.tolist() if hasattr(predictions, "tolist") else predictions
        for p, meta in zip(probs, metadata):
            self._labels.append(0 if meta["instance"][0]["label"] == "neg" else 1)
            self._scores.append(p[self._same_idx])

    def get_metric(self, reset: bool):
        out = {"precision": 0, "recall": 0, "f1": 0, "thres": 0, "auc": 0, "ave_precision_score": 0}
        if not self._scores:
            return out
        if reset:
            out = find_best_thres(self._labels, self._scores, interval=(0.5, 0.9))
            if len(set(self._labels)) > 1:
                out["auc"] = roc_auc(self._labels, self._scores)
                out["ave_precision_score"] = average_precision(self._labels, self._scores)
            else:       # sklearn is undefined with one class; the reference would emit nan + a warning
                out["auc"] = float("nan")
                out["ave_precision_score"] = float("nan")
            self.reset()
        return out

    def reset(self) -> None:
        self._labels.clear()
        self._scores.clear()


class CategoricalAccuracy:
    def __init__(self) -> None:
        self.correct = 0
        self.total = 0

    def __call__(self, predictions: np.ndarray, gold_labels: np.ndarray) -> None:
        pred = np.argmax(np.asarray(predictions), axis=-1)
        gold = np.asarray(gold_labels)
        self.correct += int(np.sum(pred == gold))
        self.total += int(gold.size)

    def get_metric(self, reset: bool = False) -> float:
        acc = self.correct / self.total if self.total else 0.0
        if reset:
            self.correct = self.total = 0
        return acc


class FBetaMeasure:
    """beta=1; ``average`` None (per-class lists) or "weighted" (support-weighted floats)."""

    def __init__(self, num_classes: int, average: Optional[str] = None) -> None:
        self.n = num_classes
        self.average = average
        self.reset()

    def reset(self) -> None:
        self.tp = np.zeros(self.n); self.pred = np.zeros(self.n); self.true = np.zeros(self.n)

    def __call__(self, predictions: np.ndarray, gold_labels: np.ndarray) -> None:
        pred = np.argmax(np.asarray(predictions), axis=-1)
        gold = np.asarray(gold_labels)
        for c in range(self.n):
            self.tp[c] += np.sum((pred == c) & (gold == c))
            self.pred[c] += np.sum(pred == c)
            self.true[c] += np.sum(gold == c)

    def get_metric(self, reset: bool = False) -> Dict[str, Any]:
        with np.errstate(divide="ignore", invalid="ignore"):
            p = np.where(self.pred > 0, self.tp / self.pred, 0.0)
            r = np.where(self.true > 0, self.tp / self.true, 0.0)
            f = np.where(p + r > 0, 2 * p * r / (p + r), 0.0)
        if self.average == "weighted":
            w = self.true / self.true.sum() if self.true.sum() else np.zeros(self.n)
            out = {"precision": float((p * w).sum()), "recall": float((r * w).sum()), "fscore": float((f * w).sum())}
        else:
            out = {"precision": p.tolist(), "recall": r.tolist(), "fscore": f.tolist()}
        if reset:
            self.reset()
        return out


class ClassificationReport:
    """The accuracy / weighted-F1 / per-class-F1 trio both models keep (model_memory.py:80-84, model_single.py:68-72) and
    the flat dict their ``get_metrics`` report (model_memory.py:196-207): one place instead of two copies."""

    def __init__(self, num_classes: int, idx2label: Dict[int, str]) -> None:
        self.idx2label = idx2label
        self.parts = {"accuracy": CategoricalAccuracy(),
                      "f1-score_overall": FBetaMeasure(num_classes, average="weighted"),
                      "f1-score_each": FBetaMeasure(num_classes, average=None)}

    def update(self, predictions, gold_labels) -> None:
        for m in self.parts.values():
            m(predictions, gold_labels)

    def report(self, reset: bool) -> Dict[str, float]:
        out: Dict[str, float] = {"accuracy": self.parts["accuracy"].get_metric(reset)}
        overall = self.parts["f1-score_overall"].get_metric(reset)
        out.update({"precision": overall["precision"], "recall": overall["recall"], "f1-score": overall["fscore"]})
        each = self.parts["f1-score_each"].get_metric(reset)
        for i, name in sorted(self.idx2label.items()):
            for short, key in (("precision", "precision"), ("recall", "recall"), ("f1-score", "fscore")):
                out[f"{name}_{short}"] = each[key][i]
        return out
