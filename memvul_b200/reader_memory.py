"""``reader_memory`` dataset reader, evaluation branches.

Mirrors MemVul/reader_memory.py:35-246 for the paths ``predict_memory.py`` exercises:
  * file-name dispatch on the substrings ``golden_`` / ``test_`` / ``validation_`` (:138,146,155);
  * anchors: ``{cwe_id: description}`` JSON -> one ``golden`` instance each (:73-79,138-144);
  * issue reports: text = ``"{Issue_Title}. {Issue_Body}"`` (:88), positives keyed by the CWE id of their
    CVE (:92-105, samples whose CWE id is None are dropped), emitted in REVERSED concatenation order so
    positives come first (:150,158); ``type`` is ``unlabel`` for test files and ``test`` for validation files;
  * fields: ``sample1`` (word-piece ids + type ids), ``label`` ``same``/``diff`` (:231-237), ``metadata``
    ``{"type", "instance": [{"label", "Issue_Url"}]}`` (:239-245).
Pair sampling for training (:164-192, 203-224) is out of scope and raises.
An instance is a plain dict (see memvul_b200/collate.py) instead of an AllenNLP ``Instance``.
"""
from __future__ import annotations

import json
import logging
from typing import Any, Dict, Iterator, List, Optional

from .registrable import DatasetReader
from .tokenizer import build_tokenizer

logger = logging.getLogger(__name__)


@DatasetReader.register("reader_memory")
class ReaderMemory(DatasetReader):
    def __init__(self,
                 tokenizer=None,
                 same_diff_ratio: Dict[str, int] = None,
                 target: str = "Security_Issue_Full",
                 anchor_path: str = "CWE_anchor_golden_project.json",
                 sample_neg: float = None,
                 train_iter: int = None,
                 token_indexers: Dict[str, Any] = None,
                 cve_dict_path: Optional[str] = None,
                 label_vocab: Optional[Dict[str, int]] = None) -> None:
        super().__init__()
        self._tokenizer = build_tokenizer(tokenizer) if tokenizer is not None else None
        self._token_indexers = token_indexers
        self._same_diff_ratio = same_diff_ratio or {"diff": 6, "same": 2}
        self._target = target
        self._anchor_path = anchor_path
        self._sample_neg = sample_neg
        self._train_iter = train_iter or 1
        self._label_vocab = label_vocab
        self._cve_info: Dict[str, Any] = {}
        if cve_dict_path:
            with open(cve_dict_path, encoding="utf-8") as f:
                self._cve_info = json.load(f)
        self._dataset: Dict[str, Dict[str, list]] = {}

    def index_with(self, vocab, namespace: str = "labels") -> None:
        self._label_vocab = {t: vocab.get_token_index(t, namespace) for t in ("same", "diff")}

    # ------------------------------------------------------------------ reading
    TOKENIZE_CHUNK = 512      # texts per tokenize_batch call: the Rust backend encodes a chunk on all cores, GIL released

    def _tokenize_all(self, texts: List[str]) -> List[List[str]]:
        if hasattr(self._tokenizer, "tokenize_batch"):
            return self._tokenizer.tokenize_batch(texts)
        return [self._tokenizer.tokenize(t) for t in texts]

    def _group(self, file_path: str) -> Dict[str, list]:
        """reader_memory.py:82-111 without the tokenisation: label every sample, key positives by the CWE id of their
        CVE, drop positives whose CWE id is None.  Groups keep first-seen order (``neg`` first, :84)."""
        with open(file_path, encoding="utf-8") as f:
            samples = json.load(f)
        dataset: Dict[str, list] = {"neg": []}
        for s in samples:
            label = "pos" if str(s[self._target]) == "1" else "neg"
            s[self._target] = label
            if label == "pos":
                s["CWE_ID"] = self._cve_info[s["CVE_ID"]]["CWE_ID"] if self._cve_info else s.get("CWE_ID")
                label = s["CWE_ID"]
                if label is None:
                    continue
                dataset.setdefault(label, [])
            dataset[label].append(s)
        return dataset

    def read_dataset(self, file_path: str) -> Dict[str, list]:
        if "golden" in file_path:
            with open(file_path, encoding="utf-8") as f:
                anchors = json.load(f)
            toks = self._tokenize_all(list(anchors.values()))
            return {cwe: [{self._target: cwe, "description": t}] for cwe, t in zip(anchors, toks)}
        if file_path in self._dataset:
            return self._dataset[file_path]
        dataset = self._group(file_path)
        flat = [s for group in dataset.values() for s in group]
        for s, t in zip(flat, self._tokenize_all([f"{s['Issue_Title']}. {s['Issue_Body']}" for s in flat])):
            s["description"] = t
        self._dataset[file_path] = dataset
        return dataset

    def _read(self, file_path: str) -> Iterator[Dict[str, Any]]:
        if "golden_" in file_path:
            for group in self.read_dataset(file_path).values():
                for sample in group:
                    yield self.text_to_instance((sample, sample), type_="golden")
            return
        if "test_" in file_path:
            type_ = "unlabel"
        elif "validation_" in file_path:
            type_ = "test"
        else:
            raise NotImplementedError("training-pair sampling (reader_memory.py:164-192) is out of scope for "
                                      "memvul_b200; file names must contain golden_, test_ or validation_")
        # Evaluation streams are emitted in REVERSED concatenation order (:150,158) and tokenised lazily, chunk by chunk,
        # so that a consumer (predict_memory.evaluate's prefetch thread) overlaps tokenisation with the GPU.
        dataset = self._dataset.get(file_path) or self._group(file_path)
        order = [s for group in dataset.values() for s in group][::-1]
        for c0 in range(0, len(order), self.TOKENIZE_CHUNK):
            chunk = order[c0:c0 + self.TOKENIZE_CHUNK]
            todo = [s for s in chunk if "description" not in s]
            for s, t in zip(todo, self._tokenize_all([f"{s['Issue_Title']}. {s['Issue_Body']}" for s in todo])):
                s["description"] = t
            for sample in chunk:
                yield self.text_to_instance((sample, sample), type_=type_)
        self._dataset[file_path] = dataset

    def text_to_instance(self, p, type_: str = "train") -> Dict[str, Any]:
        ins1, _ = p
        if type_ == "train":
            raise NotImplementedError("training pairs are out of scope for memvul_b200")
        tokens = ins1["description"]
        ids = self._tokenizer.ids(tokens)
        inst: Dict[str, Any] = {"sample1": {"token_ids": ids, "type_ids": [0] * len(ids)}, "label": None}
        cls1 = ins1[self._target]
        if type_ in ("test", "unlabel"):
            name = "same" if cls1 == "pos" else "diff"
            inst["label_str"] = name
            if self._label_vocab is not None:
                inst["label"] = self._label_vocab[name]
        meta = {"label": cls1}
        if type_ in ("test", "unlabel"):
            if cls1 == "pos":
                meta["label"] = ins1["CWE_ID"]
            meta["Issue_Url"] = ins1["Issue_Url"]
        inst["metadata"] = {"type": type_, "instance": [meta]}
        return inst
