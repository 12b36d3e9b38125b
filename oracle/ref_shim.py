"""TEST INFRASTRUCTURE -- stand-ins for the third-party packages the reference imports but this image lacks.

Purpose: let `oracle/make_reference_golden.py` import and EXECUTE the reference's own, unmodified first-party files
(`/root/reference/MemVul/model_memory.py`, `custom_PTM_embedder.py`, `custom_metric.py`, `/root/reference/predict_memory.py`)
in this container, so that the golden vectors under `tests/golden/ref_*` are outputs of the reference's code and not of
a restatement.  Nothing here is imported by the product, and nothing here travels as a dependency of a test: the tests
read the committed fixtures only.

What is real and what is a stand-in during such a run:
  * real, unmodified: every line of the reference files named above (ModelMemory.__init__/_instance_forward/
    forward_gold_instances/forward/make_output_human_readable/get_metrics, the embedder's forward, SiameseMeasureV1,
    find_best_thres, cal_metrics, model_measure);
  * real but a different version: `transformers.BertModel` (5.5 here, the reference pins 4.1.0 -- same arithmetic for
    rows with at least one unmasked key), sklearn, numpy, torch;
  * stand-ins written here from AllenNLP 2.4.0's documented behaviour (the package is absent):
    `Registrable.register`, `Model`, `Vocabulary`, `BasicTextFieldEmbedder` (kwargs routed by the embedder's forward
    signature), `BertPooler` (= dropout(HF BertPooler(tokens))), `FeedForward` (= dropout(act(linear(x))) per layer),
    `CategoricalAccuracy`, `FBetaMeasure`, `PretrainedTransformerTokenizer` (special-token counts; WordPiece via the
    `tokenizers` library when given a vocab file), `PretrainedTransformerIndexer`, `DatasetReader.read`, `TextField` /
    `LabelField` / `MetadataField` / `Instance` (plain holders),
    `InitializerApplicator` (no-op), and `overrides` (identity decorator).  Every other name the reference merely
    imports (`matplotlib`, `spacy`, trainer classes, ...) resolves to an inert placeholder.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import inspect
import sys
import types
from typing import Any, Dict, List, Optional

import torch

ROOTS = ("allennlp", "overrides", "matplotlib", "spacy", "_jsonnet")
SETTINGS = {"hidden": 768, "vocab_size": 30522}      # what AllenNLP would read from the hub model "bert-base-uncased"


# ----------------------------------------------------------------------------- AllenNLP behaviour, restated
class Registrable:
    _names: Dict[str, type] = {}

    @classmethod
    def register(cls, name: str, constructor: Optional[str] = None, exist_ok: bool = False):
        def add(sub: type) -> type:
            Registrable._names[name] = sub
            return sub
        return add


class Vocabulary:
    def __init__(self, namespaces: Dict[str, List[str]]) -> None:
        self._t2i = {ns: {t: i for i, t in enumerate(toks)} for ns, toks in namespaces.items()}

    def get_token_index(self, token: str, namespace: str = "tokens") -> int:
        return self._t2i[namespace][token]

    def get_index_to_token_vocabulary(self, namespace: str = "tokens") -> Dict[int, str]:
        return {i: t for t, i in self._t2i[namespace].items()}

    def get_vocab_size(self, namespace: str = "tokens") -> int:
        return len(self._t2i[namespace])


class Model(torch.nn.Module, Registrable):
    def __init__(self, vocab: Vocabulary, regularizer: Any = None) -> None:
        super().__init__()
        self.vocab = vocab
        self._regularizer = regularizer


class TokenEmbedder(torch.nn.Module, Registrable):
    pass


class TextFieldEmbedder(torch.nn.Module, Registrable):
    pass


class BasicTextFieldEmbedder(TextFieldEmbedder):
    """One sub-module `token_embedder_<key>` per indexer key; each receives the tensors of its key whose names appear
    in its forward signature; outputs are concatenated on the last dimension."""

    def __init__(self, token_embedders: Dict[str, torch.nn.Module]) -> None:
        super().__init__()
        self._keys = sorted(token_embedders)
        for k in self._keys:
            self.add_module("token_embedder_" + k, token_embedders[k])

    def get_output_dim(self) -> int:
        return sum(getattr(self, "token_embedder_" + k).get_output_dim() for k in self._keys)

    def forward(self, text_field_input: Dict[str, Dict[str, torch.Tensor]], num_wrapping_dims: int = 0, **kwargs):
        outs = []
        for k in self._keys:
            emb = getattr(self, "token_embedder_" + k)
            names = set(inspect.signature(emb.forward).parameters)
            outs.append(emb(**{n: t for n, t in text_field_input[k].items() if n in names}))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=-1)


class BertPooler(torch.nn.Module, Registrable):
    """`dropout(pooler(tokens))` with HF's BertPooler (dense + tanh on token 0).  AllenNLP deep-copies the pooler of
    the named hub model; here it is freshly initialised at the configured width and then overwritten by load_state_dict."""

    def __init__(self, pretrained_model: str, *, requires_grad: bool = True, dropout: float = 0.0, **kwargs: Any) -> None:
        super().__init__()
        from transformers import BertConfig
        from transformers.models.bert.modeling_bert import BertPooler as HFBertPooler
        self.pooler = HFBertPooler(BertConfig(hidden_size=SETTINGS["hidden"]))
        for p in self.pooler.parameters():
            p.requires_grad = requires_grad
        self._dropout = torch.nn.Dropout(p=dropout)
        self._embedding_dim = SETTINGS["hidden"]

    def get_input_dim(self) -> int:
        return self._embedding_dim

    def get_output_dim(self) -> int:
        return self._embedding_dim

    def forward(self, tokens: torch.Tensor, mask: torch.BoolTensor = None, num_wrapping_dims: int = 0):
        return self._dropout(self.pooler(tokens))


class FeedForward(torch.nn.Module, Registrable):
    def __init__(self, input_dim: int, num_layers: int, hidden_dims, activations, dropout=0.0) -> None:
        super().__init__()
        as_list = lambda v: list(v) if isinstance(v, (list, tuple)) else [v] * num_layers      # noqa: E731
        hidden_dims, activations, dropout = as_list(hidden_dims), as_list(activations), as_list(dropout)
        dims = [input_dim] + hidden_dims
        self._activations = torch.nn.ModuleList(activations)
        self._linear_layers = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self._dropout = torch.nn.ModuleList([torch.nn.Dropout(p=v) for v in dropout])
        self._output_dim = hidden_dims[-1]
        self.input_dim = input_dim

    def get_output_dim(self) -> int:
        return self._output_dim

    def get_input_dim(self) -> int:
        return self.input_dim

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        out = inputs
        for layer, act, drop in zip(self._linear_layers, self._activations, self._dropout):
            out = drop(act(layer(out)))
        return out


class InitializerApplicator:
    def __init__(self, *a: Any, **k: Any) -> None:
        pass

    def __call__(self, module: torch.nn.Module) -> None:
        pass


class RegularizerApplicator:
    pass


class Metric(Registrable):
    pass


class CategoricalAccuracy(Metric):
    def __init__(self, top_k: int = 1, tie_break: bool = False) -> None:
        self.correct_count, self.total_count = 0.0, 0.0

    def __call__(self, predictions: torch.Tensor, gold_labels: torch.Tensor, mask=None) -> None:
        predictions, gold_labels = predictions.detach(), gold_labels.detach()
        top1 = predictions.max(-1)[1]
        self.correct_count += float((top1 == gold_labels.long()).sum())
        self.total_count += float(gold_labels.numel())

    def get_metric(self, reset: bool = False) -> float:
        acc = self.correct_count / self.total_count if self.total_count > 0 else 0.0
        if reset:
            self.correct_count, self.total_count = 0.0, 0.0
        return acc


class FBetaMeasure(Metric):
    def __init__(self, beta: float = 1.0, average: Optional[str] = None, labels=None) -> None:
        self._beta, self._average, self._labels = beta, average, None if labels is None else list(labels)
        self._tp = self._pred = self._true = None

    def __call__(self, predictions: torch.Tensor, gold_labels: torch.Tensor, mask=None) -> None:
        predictions, gold = predictions.detach(), gold_labels.detach().long()
        n = predictions.size(-1)
        if self._tp is None:
            self._tp, self._pred, self._true = torch.zeros(n), torch.zeros(n), torch.zeros(n)
        arg = predictions.max(-1)[1]
        self._tp += torch.bincount(gold[arg == gold], minlength=n).float()
        self._pred += torch.bincount(arg, minlength=n).float()
        self._true += torch.bincount(gold, minlength=n).float()

    def get_metric(self, reset: bool = False) -> Dict[str, Any]:
        if self._tp is None:
            raise RuntimeError("You never call this metric before.")
        tp, pred, true = self._tp, self._pred, self._true
        if self._labels is not None:
            tp, pred, true = tp[self._labels], pred[self._labels], true[self._labels]
        b2 = self._beta ** 2
        div = lambda a, b: torch.where(b == 0, torch.zeros_like(a), a / torch.where(b == 0, torch.ones_like(b), b))   # noqa: E731
        precision, recall = div(tp, pred), div(tp, true)
        fscore = div((1 + b2) * precision * recall, b2 * precision + recall)
        if self._average == "weighted":
            w = true / true.sum()
            precision, recall, fscore = (precision * w).sum(), (recall * w).sum(), (fscore * w).sum()
        if reset:
            self._tp = self._pred = self._true = None
        if self._average is None:
            return {"precision": precision.tolist(), "recall": recall.tolist(), "fscore": fscore.tolist()}
        return {"precision": precision.item(), "recall": recall.item(), "fscore": fscore.item()}


class _HFTokenizerStub:
    def __len__(self) -> int:
        return SETTINGS["vocab_size"]


class Token:
    def __init__(self, text: str, text_id: int, type_id: int = 0) -> None:
        self.text, self.text_id, self.type_id = text, text_id, type_id


class Tokenizer(Registrable):
    pass


class PretrainedTransformerTokenizer(Tokenizer):
    """With a hub name: only what the embedder's constructor reads (vocabulary size, [CLS] / [SEP] counts).
    With a path to a vocab.txt: BERT WordPiece through the `tokenizers` library -- the backend of the HF fast
    tokenizer AllenNLP wraps -- lower-cased, [CLS] ... [SEP] added, truncated to max_length."""

    def __init__(self, model_name: str, add_special_tokens: bool = True, max_length: Optional[int] = None,
                 tokenizer_kwargs: Optional[Dict[str, Any]] = None, **k: Any) -> None:
        import os
        self.tokenizer = _HFTokenizerStub()
        self.single_sequence_start_tokens = ["[CLS]"]
        self.single_sequence_end_tokens = ["[SEP]"]
        self._wp = None
        if os.path.isfile(model_name):
            import tokenizers
            self._wp = tokenizers.BertWordPieceTokenizer(model_name, lowercase=True)
            if max_length is not None:
                self._wp.enable_truncation(max_length=max_length)
            self._special = add_special_tokens

    def tokenize(self, text: str) -> List[Token]:
        enc = self._wp.encode(text, add_special_tokens=self._special)
        return [Token(t, i, ty) for t, i, ty in zip(enc.tokens, enc.ids, enc.type_ids)]


class PretrainedTransformerIndexer(Registrable):
    def __init__(self, model_name: str, namespace: str = "tags", max_length: Optional[int] = None, **k: Any) -> None:
        self._namespace = namespace

    def tokens_to_indices(self, tokens: List[Token], vocabulary: Any = None) -> Dict[str, List[Any]]:
        return {"token_ids": [t.text_id for t in tokens], "mask": [True] * len(tokens), "type_ids": [t.type_id for t in tokens]}


class TextField:
    def __init__(self, tokens: List[Token], token_indexers: Dict[str, Any] = None) -> None:
        self.tokens, self._token_indexers = tokens, token_indexers


class LabelField:
    def __init__(self, label: Any, label_namespace: str = "labels", skip_indexing: bool = False) -> None:
        self.label, self._label_namespace = label, label_namespace


class MetadataField:
    def __init__(self, metadata: Any) -> None:
        self.metadata = metadata


class Instance:
    def __init__(self, fields: Dict[str, Any]) -> None:
        self.fields = fields


class DatasetReader(Registrable):
    def __init__(self, *a: Any, **k: Any) -> None:
        pass

    def read(self, file_path: str):
        yield from self._read(file_path)


REAL: Dict[str, Dict[str, Any]] = {
    "overrides": {"overrides": lambda f: f},
    "allennlp.common": {"Registrable": Registrable},
    "allennlp.data": {"Vocabulary": Vocabulary, "TextFieldTensors": Dict[str, Dict[str, torch.Tensor]]},
    "allennlp.models": {"Model": Model},
    "allennlp.modules": {"TextFieldEmbedder": TextFieldEmbedder, "FeedForward": FeedForward},
    "allennlp.modules.text_field_embedders": {"BasicTextFieldEmbedder": BasicTextFieldEmbedder},
    "allennlp.modules.token_embedders.token_embedder": {"TokenEmbedder": TokenEmbedder},
    "allennlp.modules.seq2vec_encoders": {"BertPooler": BertPooler},
    "allennlp.nn": {"InitializerApplicator": InitializerApplicator, "RegularizerApplicator": RegularizerApplicator},
    "allennlp.training.metrics": {"Metric": Metric, "CategoricalAccuracy": CategoricalAccuracy, "FBetaMeasure": FBetaMeasure},
    "allennlp.data.tokenizers": {"PretrainedTransformerTokenizer": PretrainedTransformerTokenizer, "Tokenizer": Tokenizer},
    "allennlp.data.token_indexers": {"PretrainedTransformerIndexer": PretrainedTransformerIndexer},
    "allennlp.data.dataset_readers.dataset_reader": {"DatasetReader": DatasetReader},
    "allennlp.data.fields": {"TextField": TextField, "LabelField": LabelField, "MetadataField": MetadataField},
    "allennlp.data.instance": {"Instance": Instance},
}


# ----------------------------------------------------------------------------- import machinery
class _Placeholder(types.ModuleType):
    """Module whose unknown attributes are inert classes (the reference imports far more than the path executes)."""

    def __getattr__(self, name: str):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        cls = type(name, (Registrable,), {"__init__": lambda self, *a, **k: None, "__module__": self.__name__})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Placeholder(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module) -> None:
        for k, v in REAL.get(module.__name__, {}).items():
            setattr(module, k, v)


def install(hidden: int = 768, vocab_size: int = 30522) -> None:
    for root in ROOTS:
        try:
            importlib.import_module(root)
            if root == "allennlp":
                raise RuntimeError("a real allennlp is importable: use it instead of oracle/ref_shim.py")
        except ImportError:
            pass
    SETTINGS["hidden"], SETTINGS["vocab_size"] = hidden, vocab_size
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    # reader_memory.py imports a class transformers no longer ships (unused by the reader)
    import transformers.utils.dummy_pt_objects as dummy
    if not hasattr(dummy, "ElectraForMaskedLM"):
        dummy.ElectraForMaskedLM = type("ElectraForMaskedLM", (), {})
    # numpy 2.x dropped the private paths predict_memory.py imports from (numpy.core.*, numpy.lib.npyio)
    import numpy as np
    for name, attrs in (("numpy.core.defchararray", {"encode": np.char.encode}), ("numpy.core.fromnumeric", {"sort": np.sort}),
                        ("numpy.lib.npyio", {"load": np.load})):
        try:
            mod = importlib.import_module(name)
            for a, v in attrs.items():
                if not hasattr(mod, a):
                    setattr(mod, a, v)
        except Exception:                                     # noqa: BLE001
            mod = types.ModuleType(name)
            for a, v in attrs.items():
                setattr(mod, a, v)
            sys.modules[name] = mod


def import_reference(root: str = "/root/reference"):
    """The reference's `MemVul` directory has no __init__.py and this repo ships an alias package of the same name, so
    bind the name to the reference directory explicitly before importing its modules."""
    if "MemVul" in sys.modules and getattr(sys.modules["MemVul"], "__path__", [None])[0] != root + "/MemVul":
        raise RuntimeError("the repo's MemVul alias package is already imported in this process")
    pkg = types.ModuleType("MemVul")
    pkg.__path__ = [root + "/MemVul"]
    sys.modules["MemVul"] = pkg
    if root not in sys.path:
        sys.path.insert(0, root)
    mm = importlib.import_module("MemVul.model_memory")
    emb = importlib.import_module("MemVul.custom_PTM_embedder")
    met = importlib.import_module("MemVul.custom_metric")
    drv = importlib.import_module("predict_memory")
    if not drv.__file__.startswith(root):
        raise RuntimeError("predict_memory resolved to %s, not the reference" % drv.__file__)
    return mm, emb, met, drv
