"""AllenNLP plug-in surface for the drop-in.

The reference registers its classes with AllenNLP's ``Registrable`` decorators
(``@Model.register("model_memory")`` MemVul/model_memory.py:39,
``@TokenEmbedder.register("custom_pretrained_transformer")`` MemVul/custom_PTM_embedder.py:22,
``@DatasetReader.register("reader_memory")`` MemVul/reader_memory.py:35,
``@Metric.register("siamese_measure_v1")`` MemVul/custom_metric.py:55) and is built from
``config.json`` by ``FromParams``.  When AllenNLP is importable its own base classes are used, so the
classes in this package land in AllenNLP's registry under the reference's names.  AllenNLP is not in
this image, so otherwise a minimal stand-in with the same decorator / ``by_name`` / ``from_params``
behaviour is used; everything else in the package is identical either way.
"""
from __future__ import annotations

import inspect
import os
from collections import defaultdict
from typing import Any, Callable, Dict, Iterable, List, Optional, Type

import torch

try:  # pragma: no cover - AllenNLP is absent from the build image
    from allennlp.common import Registrable  # type: ignore
    from allennlp.data import DatasetReader, Vocabulary  # type: ignore
    from allennlp.models import Model  # type: ignore
    from allennlp.modules.token_embedders import TokenEmbedder  # type: ignore
    from allennlp.predictors import Predictor  # type: ignore
    from allennlp.training.metrics import Metric  # type: ignore
    from allennlp.training.callbacks.callback import TrainerCallback  # type: ignore
    HAVE_ALLENNLP = True
except Exception:  # noqa: BLE001
    HAVE_ALLENNLP = False

    class Registrable:
        """``cls.register(name)`` decorator + ``cls.by_name(name)`` lookup, per base class."""
        _registry: Dict[type, Dict[str, type]] = defaultdict(dict)

        @classmethod
        def register(cls, name: str, constructor: Optional[str] = None, exist_ok: bool = False) -> Callable:
            registry = Registrable._registry[cls]

            def add(subclass: type) -> type:
                if name in registry and not exist_ok and registry[name] is not subclass:
                    raise ValueError(f"Cannot register {name} as {cls.__name__}; name already in use for "
                                     f"{registry[name].__name__}")
                registry[name] = subclass
                return subclass
            return add

        @classmethod
        def by_name(cls, name: str) -> type:
            for base, reg in Registrable._registry.items():
                if issubclass(cls, base) or issubclass(base, cls):
                    if name in reg:
                        return reg[name]
            raise KeyError(f"{name} is not a registered name for {cls.__name__}")

        @classmethod
        def list_available(cls) -> List[str]:
            return sorted(Registrable._registry[cls].keys())

        @classmethod
        def from_params(cls, params: Dict[str, Any], **extras: Any):
            """Tiny FromParams: ``{"type": name, **kwargs}`` -> instance; nested dicts with a "type" whose
            parameter annotation is Registrable are built recursively."""
            params = dict(params)
            target = cls.by_name(params.pop("type")) if "type" in params else cls
            sig = inspect.signature(target.__init__)
            kwargs = {}
            for k, v in {**params, **extras}.items():
                if k not in sig.parameters:
                    if any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values()):
                        kwargs[k] = v
                        continue
                    raise TypeError(f"{target.__name__} got an unexpected config key {k!r}")
                ann = sig.parameters[k].annotation
                if isinstance(v, dict) and inspect.isclass(ann) and issubclass(ann, Registrable):
                    sub_extras = {e: extras[e] for e in ("vocab",) if e in extras
                                  and e in inspect.signature(ann.by_name(v["type"]).__init__).parameters} \
                        if "type" in v else {}
                    v = ann.from_params(v, **sub_extras)
                kwargs[k] = v
            return target(**kwargs)

    class Vocabulary:
        """The slice of ``allennlp.data.Vocabulary`` the model uses: namespace -> token <-> index."""

        def __init__(self, tokens: Optional[Dict[str, List[str]]] = None) -> None:
            self._t2i: Dict[str, Dict[str, int]] = {}
            self._i2t: Dict[str, Dict[int, str]] = {}
            for ns, toks in (tokens or {}).items():
                self.add_tokens_to_namespace(toks, ns)

        def add_tokens_to_namespace(self, tokens: Iterable[str], namespace: str) -> None:
            t2i = self._t2i.setdefault(namespace, {})
            i2t = self._i2t.setdefault(namespace, {})
            for t in tokens:
                if t not in t2i:
                    t2i[t] = len(t2i)
                    i2t[t2i[t]] = t

        @classmethod
        def from_files(cls, directory: str) -> "Vocabulary":
            """Reads ``<namespace>.txt`` files as written into a model archive's ``vocabulary/`` dir."""
            v = cls()
            for fn in sorted(os.listdir(directory)):
                if fn.endswith(".txt") and fn != "non_padded_namespaces.txt":
                    with open(os.path.join(directory, fn), encoding="utf-8") as f:
                        v.add_tokens_to_namespace([ln.rstrip("\n") for ln in f if ln.strip()], fn[:-4])
            return v

        def get_token_index(self, token: str, namespace: str = "tokens") -> int:
            return self._t2i[namespace][token]

        def get_token_from_index(self, index: int, namespace: str = "tokens") -> str:
            return self._i2t[namespace][index]

        def get_index_to_token_vocabulary(self, namespace: str = "tokens") -> Dict[int, str]:
            return dict(self._i2t.get(namespace, {}))

        def get_vocab_size(self, namespace: str = "tokens") -> int:
            return len(self._t2i.get(namespace, {}))

    class Model(torch.nn.Module, Registrable):
        def __init__(self, vocab: Vocabulary, regularizer: Any = None) -> None:
            super().__init__()
            self.vocab = vocab
            self._regularizer = regularizer

        def get_metrics(self, reset: bool = False) -> Dict[str, float]:
            return {}

        def make_output_human_readable(self, output_dict: Dict[str, Any]) -> Dict[str, Any]:
            return output_dict

        def forward_on_instances(self, instances: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
            """AllenNLP ``Model.forward_on_instances`` for the dict-instances of ``reader_memory``:
            collate (pad to the batch's longest), move to the model's device, forward in eval/no_grad."""
            from .collate import collate_instances
            device = next(self.parameters()).device
            with torch.no_grad():
                batch = collate_instances(instances, device)
                out = self.make_output_human_readable(self(**batch))
            return out if isinstance(out, list) else [out]

    class TokenEmbedder(torch.nn.Module, Registrable):
        def get_output_dim(self) -> int:
            raise NotImplementedError

    class DatasetReader(Registrable):
        def __init__(self, **kwargs: Any) -> None:
            pass

        def read(self, file_path: str):
            return self._read(file_path)

        def _read(self, file_path: str):
            raise NotImplementedError

    class Predictor(Registrable):
        def __init__(self, model: Model, dataset_reader: DatasetReader) -> None:
            self._model = model
            self._dataset_reader = dataset_reader

    class Metric(Registrable):
        def __call__(self, *args: Any, **kwargs: Any) -> None:
            raise NotImplementedError

        def get_metric(self, reset: bool):
            raise NotImplementedError

        def reset(self) -> None:
            raise NotImplementedError

    class TrainerCallback(Registrable):
        """``allennlp.training.callbacks.TrainerCallback``: only the hooks the reference's callbacks override."""

        def __init__(self, serialization_dir: Optional[str] = None) -> None:
            self.serialization_dir = serialization_dir
            self.trainer = None

        def on_start(self, trainer, is_primary: bool = True, **kwargs: Any) -> None:
            self.trainer = trainer

        def on_batch(self, trainer, *args: Any, **kwargs: Any) -> None:
            pass

        def on_epoch(self, trainer, metrics: Dict[str, Any], epoch: int, is_primary: bool = True, **kwargs: Any) -> None:
            pass

        def on_end(self, trainer, *args: Any, **kwargs: Any) -> None:
            pass


def registered(base: Type, name: str) -> type:
    return base.by_name(name)
