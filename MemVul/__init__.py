"""Alias package: the reference's configs and drivers refer to the plug-in package as ``MemVul``
(``--include-package MemVul``, ``import_module_and_submodules("MemVul")`` predict_memory.py:59).
Importing it registers the B200-native implementations under the reference's names."""
from memvul_b200 import *  # noqa: F401,F403
from memvul_b200 import (ModelMemory, ModelSingle, PretrainedTransformerEmbedder, ReaderMemory,  # noqa: F401
                         ReaderSingle, SiameseMeasureV1)
