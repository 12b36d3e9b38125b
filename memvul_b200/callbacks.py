"""Trainer callbacks of the reference that touch the hot path (MemVul/callbacks.py).

``custom_validation`` (callbacks.py:28-53) is how the reference refreshes the external memory during training: at the
end of every epoch it empties ``_golden_instances_embeddings`` / ``_golden_instances_labels`` and re-encodes the anchor
file with the current weights (first 128 anchors, then the rest, :49-52), after which the trainer's validation loop runs
``ModelMemory.forward`` with ``type == "test"`` -- the same branch predict_memory.py drives.  Here that refresh lands on
the CUDA encoder; the embedder re-packs its fp16 weight copy by itself when the fp32 parameters have changed
(custom_PTM_embedder.packed: parameter version counters), so an optimiser step between epochs is picked up.
``reset_dataloader`` (callbacks.py:16-25) only touches the training data loader and is mirrored for config compatibility.
Training itself (the backward pass) stays out of scope: see DESIGN.md.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Optional

from .predict_memory import build_memory
from .registrable import DatasetReader, TrainerCallback

logger = logging.getLogger(__name__)


@TrainerCallback.register("reset_dataloader")
class ResetLoader(TrainerCallback):
    """callbacks.py:16-25 -- the training loader re-reads (re-samples negatives) before the next epoch."""

    def on_epoch(self, trainer, metrics: Dict[str, Any], epoch: int, is_primary: bool = True, **kwargs: Any) -> None:
        trainer.data_loader._instances = None


@TrainerCallback.register("custom_validation")
class CustomValidation(TrainerCallback):
    """callbacks.py:28-53 -- rebuild the anchor bank from ``anchor_path`` with the model's current weights."""

    def __init__(self, anchor_path: str, data_reader: Optional[DatasetReader] = None, data_loader: Any = None,
                 serialization_dir: Optional[str] = None) -> None:
        super().__init__(serialization_dir)
        if data_reader is None:
            # the reference builds ReaderMemory over the hub tokenizer "bert-base-uncased" here (callbacks.py:37-39);
            # there is no hub in this image, so the vocabulary must be named explicitly
            raise ValueError("custom_validation needs data_reader (a reader_memory with a local vocab.txt)")
        self._anchors = list(data_reader.read(anchor_path))

    def on_epoch(self, trainer, metrics: Dict[str, Any], epoch: int, is_primary: bool = True, **kwargs: Any) -> None:
        model = trainer.model
        model.eval()
        model._golden_instances_embeddings = None            # reset (callbacks.py:48)
        model._golden_instances_labels = None                # reset (callbacks.py:49)
        logger.info("updating golden embeddings")
        build_memory(model, self._anchors, chunk=128)        # callbacks.py:51-53
