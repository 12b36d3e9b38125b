"""memvul_b200 -- B200-native batch inference for MemVul (BERT issue-report encoder + CWE-anchor memory match).

Importing the package registers the reference's plug-in names (``model_memory``, ``model_single``,
``reader_memory``, ``reader_single``, ``custom_pretrained_transformer``, ``siamese_measure_v1``, ``custom_validation``, ``reset_dataloader``), as
``import_module_and_submodules("MemVul")`` does for the reference (predict_memory.py:59).
"""
__version__ = "0.1.0"

from . import registrable  # noqa: F401
from .custom_metric import SiameseMeasureV1  # noqa: F401
from .custom_PTM_embedder import PretrainedTransformerEmbedder  # noqa: F401
from .model_memory import ModelMemory  # noqa: F401
from .model_single import ModelSingle  # noqa: F401
from .reader_memory import ReaderMemory  # noqa: F401
from .reader_single import ReaderSingle  # noqa: F401
from .callbacks import CustomValidation, ResetLoader  # noqa: F401
