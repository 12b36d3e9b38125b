"""``reader_single`` dataset reader (MemVul-m, BASELINE configs[0] plumbing), evaluation branches.

Mirrors MemVul/reader_single.py:30-126 for the paths ``predict_single.py`` exercises:
  * text = ``"{Issue_Title}. {Issue_Body}"`` (:60); ``pos`` iff ``str(sample[target]) == "1"`` (:62-63); samples are
    grouped by label in first-seen order (:64-66) and emitted group after group (:74-76) -- NOT reversed, unlike
    reader_memory;
  * file-name dispatch: ``test_`` -> type ``unlabel`` (:83-88), ``validation_`` -> type ``test`` (:90-94); anything
    else is the sampled training stream (:96-110), out of scope here and raises;
  * fields (:113-126): ``sample`` (word-piece ids + type ids), ``label`` in namespace ``class_labels``, ``metadata``
    ``{"type", "instance": {"Issue_Url", "label"}}`` -- ``instance`` is a dict here, a one-element list in reader_memory.
An instance is a plain dict (memvul_b200/collate.py) instead of an AllenNLP ``Instance``.
"""
from __future__ import annotations

import json
import logging
from typing import Any, Dict, Iterator, List, Optional

from .registrable import DatasetReader
from .tokenizer import build_tokenizer

logger = logging.getLogger(__name__)


@DatasetReader.register("reader_single")
class ReaderSingle(DatasetReader):
    def __init__(self,
                 tokenizer=None,
                 token_indexers: Dict[str, Any] = None,
                 sample_neg: float = None,
                 train_iter: int = None,
                 cache_directory: Optional[str] = None,
                 target: str = "Security_Issue_Full",
                 label_vocab: Optional[Dict[str, int]] = None) -> None:
        super().__init__()
        self._tokenizer = build_tokenizer(tokenizer) if tokenizer is not None else None
        self._token_indexers = token_indexers
        self._target = target
        self._train_iter = train_iter or 1
        self._sample_neg = sample_neg or 0.1
        self._label_vocab = label_vocab
        self._dataset: Dict[str, Dict[str, list]] = {}

    def index_with(self, vocab, namespace: str = "class_labels") -> None:
        self._label_vocab = {t: vocab.get_token_index(t, namespace) for t in ("pos", "neg")}

    def read_dataset(self, file_path: str) -> Dict[str, list]:
        if file_path in self._dataset:                       # tokenisation results are reused (:54-56)
            return self._dataset[file_path]
        with open(file_path, encoding="utf-8") as f:
            samples = json.load(f)
        texts = [f"{s['Issue_Title']}. {s['Issue_Body']}" for s in samples]
        toks = self._tokenizer.tokenize_batch(texts) if hasattr(self._tokenizer, "tokenize_batch") \
            else [self._tokenizer.tokenize(t) for t in texts]
        dataset: Dict[str, list] = {}
        for s, t in zip(samples, toks):
            s["description"] = t
            label = "pos" if str(s[self._target]) == "1" else "neg"
            s[self._target] = label
            dataset.setdefault(label, []).append(s)
        self._dataset[file_path] = dataset
        return dataset

    def _read(self, file_path: str) -> Iterator[Dict[str, Any]]:
        dataset = self.read_dataset(file_path)
        all_data: List[dict] = []
        for group in dataset.values():
            all_data.extend(group)
        logger.info({k: len(v) for k, v in dataset.items()})
        if "test_" in file_path:
            for sample in all_data:
                yield self.text_to_instance(sample, type_="unlabel")
        elif "validation_" in file_path:
            for sample in all_data:
                yield self.text_to_instance(sample, type_="test")
        else:
            raise NotImplementedError("the sampled training stream (reader_single.py:96-110) is out of scope for "
                                      "memvul_b200; file names must contain test_ or validation_")

    def text_to_instance(self, ins, type_: str = "train") -> Dict[str, Any]:
        ids = self._tokenizer.ids(ins["description"])
        name = ins[self._target]
        inst: Dict[str, Any] = {"sample": {"token_ids": ids, "type_ids": [0] * len(ids)}, "label_str": name,
                                "label": None if self._label_vocab is None else self._label_vocab[name]}
        inst["metadata"] = {"type": type_, "instance": {"Issue_Url": ins["Issue_Url"], "label": name}}
        return inst
