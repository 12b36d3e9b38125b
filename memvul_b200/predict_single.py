"""Batch-inference driver of MemVul-m: the B200 counterpart of the reference's ``predict_single.py`` (BASELINE
configs[0]: "predict_single.py MemVul-m bert-base seq_len=128 batch=4"; the reference runs it on CPU through AllenNLP,
here the same flow runs on the GPU kernels -- there is no CPU path).

  * ``test``        -- predict_single.py:46-97: load the archive with overrides, read the evaluation file with the
                       archive's ``dataset_reader``, run every batch through ``ModelSingle.forward`` under no_grad, write
                       one JSON array of ``make_output_human_readable`` rows per batch per line (AllenNLP ``evaluate``
                       with ``predictions_output_file``), return ``get_metrics(reset=True)``.
  * ``cal_metrics`` -- predict_single.py:100-123: rows -> (label, predict, prob) -> ``model_measure``.
"""
from __future__ import annotations

import json
import logging
from typing import Any, Dict, Optional

import torch

from .collate import batches, collate_instances
from .predict_memory import load_archive, model_measure

logger = logging.getLogger(__name__)


def test(archive_file, input_file, test_config=None, weights_file=None, output_file=None, predictions_output_file=None,
         batch_size=64, cuda_device=0, seed=2021, package="MemVul", batch_weight_key="",
         file_friendly_logging=False) -> Dict[str, Any]:
    archive = load_archive(archive_file, weights_file=weights_file, cuda_device=cuda_device, overrides=test_config or {})
    model = archive.model
    model.eval()
    reader = archive.dataset_reader
    if hasattr(reader, "index_with"):
        reader.index_with(model.vocab)
    logger.info("Reading evaluation data from %s", input_file)
    loader_cfg = archive.config.get("validation_data_loader") or archive.config.get("data_loader") or {}
    bs = batch_size or loader_cfg.get("batch_size", 64)
    device = torch.device(f"cuda:{cuda_device}")
    instances = list(reader.read(input_file))
    pred_f = open(predictions_output_file, "w", encoding="utf-8") if predictions_output_file else None
    with torch.no_grad():
        for chunk in batches(instances, bs):
            batch = collate_instances(chunk, device, key="sample")
            out = model(batch["sample"], label=batch.get("label"), metadata=batch["metadata"])
            if pred_f is not None:
                pred_f.write(json.dumps(model.make_output_human_readable(out)) + "\n")
    if pred_f is not None:
        pred_f.close()
    metrics = model.get_metrics(reset=True)
    if output_file:
        with open(output_file, "w", encoding="utf-8") as f:
            json.dump(metrics, f, indent=4, default=float)
    logger.info("Finished evaluating.")
    return metrics


def cal_metrics(result_file: str, out_file: Optional[str] = None) -> Dict[str, Any]:
    """predict_single.py:100-123 with explicit paths instead of the hard-coded DATA_PATH layout."""
    rows = []
    with open(result_file, encoding="utf-8") as f:
        for line in f:
            if line.strip():
                rows.extend(json.loads(line))
    conv = {"pos": 1, "neg": 0}
    pred = [conv[r["predict"]] for r in rows]
    label = [conv[r["label"]] for r in rows]
    score = [r["prob"] for r in rows]
    metrics, _, _ = model_measure(label, pred, score, [r["Issue_Url"] for r in rows])
    if out_file:
        with open(out_file, "w", encoding="utf-8") as f:
            json.dump(metrics, f, indent=4, default=float)
    return metrics
