"""Batch-inference driver: the B200 counterpart of the reference's ``predict_memory.py``.

  * ``load_archive``  -- ``allennlp.models.archival.load_archive`` for a ``model.tar.gz`` (config.json +
                         vocabulary/ + weights.th) with dict overrides (predict_memory.py:60-67).
  * ``test_siamese``  -- predict_memory.py:49-114: build the anchor bank in chunks of 128 (:81-83), then
                         stream the evaluation data through ``ModelMemory.forward`` batch by batch, writing
                         one JSON array per batch per line (what AllenNLP ``evaluate`` does with
                         ``predictions_output_file``), and return ``model.get_metrics(reset=True)``.
  * ``cal_metrics`` / ``model_measure`` -- predict_memory.py:117-197: max over anchors -> threshold -> pos/neg,
                         confusion matrix, ROC-AUC, average precision.
Host work is overlapped with the GPU: batch i+1 is collated and enqueued before batch i's results are
turned into JSON.
"""
from __future__ import annotations

import copy
import json
import logging
import os
import tarfile
import tempfile
from typing import Any, Dict, Iterable, List, Optional

import numpy as np
import torch

from .collate import batches, collate_instances
from .custom_metric import average_precision, roc_auc
from .registrable import DatasetReader, Model, Vocabulary

logger = logging.getLogger(__name__)


def _merge(base: Dict[str, Any], over: Dict[str, Any]) -> Dict[str, Any]:
    out = copy.deepcopy(base)
    for k, v in (over or {}).items():
        out[k] = _merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else copy.deepcopy(v)
    return out


class Archive:
    def __init__(self, model, config, dataset_reader, validation_dataset_reader):
        self.model, self.config = model, config
        self.dataset_reader, self.validation_dataset_reader = dataset_reader, validation_dataset_reader


def load_archive(archive_file: str, weights_file: Optional[str] = None, cuda_device: int = -1,
                 overrides: Optional[Dict[str, Any]] = None) -> Archive:
    """``archive_file``: a ``model.tar.gz`` or an already-extracted serialization directory."""
    tmp = None
    if os.path.isdir(archive_file):
        root = archive_file
    else:
        tmp = tempfile.TemporaryDirectory()
        with tarfile.open(archive_file, "r:*") as tar:
            tar.extractall(tmp.name, filter="data")
        root = tmp.name
    with open(os.path.join(root, "config.json"), encoding="utf-8") as f:
        config = _merge(json.load(f), overrides or {})
    vocab = Vocabulary.from_files(os.path.join(root, "vocabulary"))
    model_cfg = dict(config["model"])
    model = Model.from_params(model_cfg, vocab=vocab)
    state = torch.load(weights_file or os.path.join(root, "weights.th"), map_location="cpu")
    state = {k: v for k, v in state.items() if not k.endswith("position_ids")}      # authorized_missing_keys
    missing, unexpected = model.load_state_dict(state, strict=False)
    if missing or unexpected:
        raise RuntimeError(f"archive weights do not match the model: missing={missing[:5]} unexpected={unexpected[:5]}")
    if cuda_device >= 0:
        model.cuda(cuda_device)
    reader = DatasetReader.from_params(dict(config["dataset_reader"])) if "dataset_reader" in config else None
    vreader = DatasetReader.from_params(dict(config["validation_dataset_reader"])) \
        if "validation_dataset_reader" in config else reader
    if tmp is not None:
        tmp.cleanup()
    return Archive(model, config, reader, vreader)


def build_memory(model, golden_instances: List[Dict[str, Any]], chunk: int = 128) -> None:
    """predict_memory.py:81-83: first 128 anchors, then the rest, through ``forward_on_instances``."""
    model.forward_on_instances(golden_instances[:chunk])
    if len(golden_instances) > chunk:
        model.forward_on_instances(golden_instances[chunk:])


def _prefetched(instances: Iterable[Dict[str, Any]], batch_size: int, device: torch.device, depth: int = 3):
    """Data-order batches built by a background thread: the reader's lazy tokenisation (Rust, GIL released) and the
    pinned-memory collation of batch i+1.. run while the main thread keeps the GPU fed with batch i.  Yields
    (host batch, instance count); exceptions of the producer are re-raised in the consumer."""
    import queue
    import threading
    q: "queue.Queue" = queue.Queue(maxsize=depth)
    END = object()

    def produce():
        try:
            chunk: List[Dict[str, Any]] = []
            for inst in instances:
                chunk.append(inst)
                if len(chunk) == batch_size:
                    q.put((collate_instances(chunk, device, host_only=True), len(chunk)))
                    chunk = []
            if chunk:
                q.put((collate_instances(chunk, device, host_only=True), len(chunk)))
            q.put(END)
        except BaseException as e:          # noqa: BLE001 -- handed to the consumer
            q.put(e)
    threading.Thread(target=produce, daemon=True).start()
    while True:
        item = q.get()
        if item is END:
            return
        if isinstance(item, BaseException):
            raise item
        yield item


def evaluate(model, instances: Iterable[Dict[str, Any]], batch_size: int, device: torch.device,
             predictions_output_file: Optional[str] = None, output_file: Optional[str] = None,
             bucket_by_length: bool = False) -> Dict[str, Any]:
    """AllenNLP ``evaluate`` for this model: forward every batch under no_grad, write readable predictions (one JSON
    array per ``batch_size`` instances per line, in data order).  Data order (default): instances are consumed as a
    stream through a prefetch thread, padded tokens cost nothing (packed execution).  ``bucket_by_length`` groups
    instances of similar length into the same batch (collate.plan_length_buckets) and restores data order on output; it
    needs the whole data set up front and only matters for the padded execution (MEMVUL_ENC_PACKED=0)."""
    from .collate import batch_to_device, plan_length_buckets
    pred_f = open(predictions_output_file, "w", encoding="utf-8") if predictions_output_file else None
    if not bucket_by_length:
        prev = None
        with torch.no_grad():
            for host_batch, _ in _prefetched(instances, batch_size, device):
                out = model(**batch_to_device(host_batch, device))
                if prev is not None and pred_f is not None:   # batch i-1's host work overlaps batch i's kernels
                    pred_f.write(json.dumps(model.make_output_human_readable(prev)) + "\n")
                prev = out
            if prev is not None and pred_f is not None:
                pred_f.write(json.dumps(model.make_output_human_readable(prev)) + "\n")
    else:
        instances = list(instances)
        n = len(instances)
        plan = plan_length_buckets([len(i["sample1"]["token_ids"]) for i in instances], batch_size)
        rows: List[Any] = [None] * n
        written = 0

        def drain(out, idx):
            nonlocal written
            if pred_f is None:
                return
            for i, row in zip(idx, model.make_output_human_readable(out)):
                rows[i] = row
            while written < n:                                    # emit every complete data-order line
                hi = min(n, written + batch_size)
                if any(rows[i] is None for i in range(written, hi)):
                    break
                pred_f.write(json.dumps(rows[written:hi]) + "\n")
                for i in range(written, hi):
                    rows[i] = True
                written = hi
        prev = None
        with torch.no_grad():
            for idx in plan:
                batch = collate_instances([instances[i] for i in idx], device)
                out = model(**batch)
                if prev is not None:
                    drain(*prev)
                prev = (out, idx)
            if prev is not None:
                drain(*prev)
    if pred_f:
        pred_f.close()
    metrics = model.get_metrics(reset=True)
    if output_file:
        with open(output_file, "w", encoding="utf-8") as f:
            json.dump(metrics, f, indent=4, default=float)
    return metrics


def test_siamese(archive_file, input_file, input_golden_file, test_config=None, weights_file=None, output_file=None,
                 predictions_output_file=None, batch_size=64, cuda_device=0, seed=2021, package="MemVul",
                 batch_weight_key="", file_friendly_logging=False, bucket_by_length=False) -> Dict[str, Any]:
    archive = load_archive(archive_file, weights_file=weights_file, cuda_device=cuda_device, overrides=test_config or {})
    model = archive.model
    model.eval()
    for r in (archive.dataset_reader, archive.validation_dataset_reader):
        if r is not None and hasattr(r, "index_with"):
            r.index_with(model.vocab)
    logger.info("Reading golden data from %s", input_golden_file)
    golden = list(archive.validation_dataset_reader.read(input_golden_file))
    build_memory(model, golden)
    logger.info("Reading evaluation data from %s", input_file)
    loader_cfg = archive.config.get("validation_data_loader") or archive.config.get("data_loader") or {}
    bs = batch_size or loader_cfg.get("batch_size", 64)
    device = torch.device(f"cuda:{cuda_device}")
    metrics = evaluate(model, archive.dataset_reader.read(input_file), bs, device,
                       predictions_output_file=predictions_output_file, output_file=output_file,
                       bucket_by_length=bucket_by_length)
    logger.info("Finished evaluating.")
    return metrics


def model_measure(test_label, pred, pred_score, sample_id=None):
    y = np.asarray(test_label).astype(bool)
    p = np.asarray(pred).astype(bool)
    TP, FN = int(np.sum(p & y)), int(np.sum(~p & y))
    TN, FP = int(np.sum(~p & ~y)), int(np.sum(p & ~y))
    pd = TP / (TP + FN) if TP + FN else 0
    prec = TP / (TP + FP) if TP + FP else 0
    f1 = 2 * pd * prec / (pd + prec) if pd + prec else 0
    from sklearn import metrics as skm
    fpr, tpr, _ = skm.roc_curve(test_label, pred_score, pos_label=1)
    result = {"TP": TP, "FN": FN, "TN": TN, "FP": FP, "pd&recall": pd, "prec": prec, "f1": f1,
              "ap": average_precision(test_label, pred_score), "auc": roc_auc(test_label, pred_score)}
    return result, fpr, tpr


def vote(rows: List[Dict[str, Any]], thres: float = 0.5) -> List[Dict[str, Any]]:
    """predict_memory.py:168-177: ``prob`` = max over anchors of P(same); ``predict`` = pos iff prob >= thres."""
    for s in rows:
        vote_prob = np.max(list(s["predict"].values()))
        s["prob"] = vote_prob
        s["predict"] = "pos" if vote_prob >= thres else "neg"
    return rows


def cal_metrics(result_file: str, thres: float = 0.5, out_file: Optional[str] = None) -> Dict[str, Any]:
    rows: List[Dict[str, Any]] = []
    with open(result_file, encoding="utf-8") as f:
        for line in f:
            if line.strip():
                rows.extend(json.loads(line))
    vote(rows, thres)
    conv = {"pos": 1, "neg": 0}
    pred = [conv[r["predict"]] for r in rows]
    label = [0 if r["label"] == "neg" else 1 for r in rows]
    score = [r["prob"] for r in rows]
    metrics, _, _ = model_measure(label, pred, score, [r["Issue_Url"] for r in rows])
    metrics["thres"] = thres
    if out_file:
        with open(out_file, "w", encoding="utf-8") as f:
            json.dump(metrics, f, indent=4, default=float)
    return metrics
