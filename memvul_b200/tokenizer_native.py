"""ctypes binding of ``libmemvul_tok.so`` (C ABI in ``include/memvul_tok.h``): the native batched WordPiece tokenizer
of the front-end (the reference tokenises with the HF fast tokenizer through AllenNLP,
config_memory.json:12-20, reader_memory.py:76,88).

``NativeWordPieceTokenizer.encode_batch`` turns a list of texts into the padded int64 id matrix + lengths in one call
(all host cores, GIL released, no per-token Python objects).  Texts the native ASCII path does not cover (non-ASCII
bytes, literal special tokens) go through the Unicode-complete fallback (HF ``tokenizers`` if importable, else the
pure-Python ``WordPieceTokenizer``) -- same results either way.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libmemvul_tok.so")
EXPORTS = ["memvul_tok_create", "memvul_tok_destroy", "memvul_tok_last_error", "memvul_tok_token_to_id",
           "memvul_tok_encode_batch"]
_lock = threading.Lock()
_lib: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "csrc", "wordpiece.cpp"), os.path.join(_REPO, "include", "memvul_tok.h")]
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", LIB_PATH, srcs[0]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                build()
            L = ctypes.CDLL(LIB_PATH)
            L.memvul_tok_create.restype = ctypes.c_void_p
            L.memvul_tok_create.argtypes = [ctypes.c_char_p, ctypes.c_int]
            L.memvul_tok_destroy.argtypes = [ctypes.c_void_p]
            L.memvul_tok_last_error.restype = ctypes.c_char_p
            L.memvul_tok_token_to_id.restype = ctypes.c_int32
            L.memvul_tok_token_to_id.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
            L.memvul_tok_encode_batch.restype = ctypes.c_int
            L.memvul_tok_encode_batch.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
            _lib = L
        return _lib


class EncodedText:
    """Word pieces of one text: the ids as a numpy view; the piece strings only on demand."""
    __slots__ = ("ids_np", "_itos")

    def __init__(self, ids_np: np.ndarray, itos: Sequence[str]) -> None:
        self.ids_np, self._itos = ids_np, itos

    @property
    def ids(self) -> List[int]:
        return self.ids_np.tolist()

    def _tokens(self) -> List[str]:
        return [self._itos[i] for i in self.ids_np.tolist()]

    def __len__(self) -> int:
        return int(self.ids_np.shape[0])

    def __iter__(self):
        return iter(self._tokens())

    def __getitem__(self, i):
        return self._tokens()[i]

    def __eq__(self, other) -> bool:
        return list(self) == list(other)


class NativeWordPieceTokenizer:
    NO_LIMIT = 4096           # row width when max_length is None (BERT itself stops at 512 positions)

    def __init__(self, vocab_file: str, lowercase: bool = True, add_special_tokens: bool = True,
                 max_length: Optional[int] = None, fallback=None, n_threads: int = 0) -> None:
        self._h = lib().memvul_tok_create(vocab_file.encode(), 1 if lowercase else 0)
        if not self._h:
            raise RuntimeError("memvul_tok_create failed: " + lib().memvul_tok_last_error().decode("utf-8", "replace"))
        with open(vocab_file, encoding="utf-8") as f:
            self._itos = [ln.rstrip("\n") for ln in f]
        self.vocab = {t: i for i, t in enumerate(self._itos) if t != ""}
        self.lowercase, self.add_special, self.max_length, self.n_threads = lowercase, add_special_tokens, max_length, n_threads
        self.unk, self.cls, self.sep = "[UNK]", "[CLS]", "[SEP]"
        if fallback is None:
            from .tokenizer import FastWordPieceTokenizer, WordPieceTokenizer
            try:
                fallback = FastWordPieceTokenizer(vocab_file, lowercase, add_special_tokens, max_length)
            except ImportError:
                fallback = WordPieceTokenizer(vocab_file, lowercase, add_special_tokens, max_length)
        self._fallback = fallback
        self.fallback_count = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().memvul_tok_destroy(self._h)
        except Exception:       # noqa: BLE001 -- interpreter shutdown
            pass

    def encode_batch(self, texts: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        """-> (ids int64 [n, W] zero-padded, lens int32 [n]); W = max_length (or the longest row when unlimited)."""
        n = len(texts)
        W = self.max_length or self.NO_LIMIT
        raw = [t.encode("utf-8") for t in texts]
        offs = np.zeros(n + 1, dtype=np.int64)
        if n:
            np.cumsum([len(r) for r in raw], out=offs[1:])
        data = b"".join(raw)
        ids = np.zeros((n, W), dtype=np.int64)
        lens = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint8)
        rc = lib().memvul_tok_encode_batch(self._h, data, offs.ctypes.data, n, 1 if self.add_special else 0, W,
                                           ids.ctypes.data, lens.ctypes.data, status.ctypes.data, self.n_threads)
        if rc < 0:
            raise ValueError(lib().memvul_tok_last_error().decode("utf-8", "replace"))
        if rc > 0:                                      # Unicode / special-token texts: the complete tokenizer
            idx = np.nonzero(status)[0].tolist()
            self.fallback_count += len(idx)
            toks = self._fallback.tokenize_batch([texts[i] for i in idx]) if hasattr(self._fallback, "tokenize_batch") \
                else [self._fallback.tokenize(texts[i]) for i in idx]
            for i, t in zip(idx, toks):
                row = self._fallback.ids(t)[:W]
                ids[i, :len(row)] = row
                lens[i] = len(row)
        if self.max_length is None and n:
            ids = ids[:, :max(int(lens.max()), 1)]
        return ids, lens

    # ---- the tokenizer interface the readers use ----
    def tokenize_batch(self, texts: Sequence[str]) -> List[EncodedText]:
        ids, lens = self.encode_batch(texts)
        return [EncodedText(ids[i, :lens[i]], self._itos) for i in range(len(texts))]

    def tokenize(self, text: str) -> EncodedText:
        return self.tokenize_batch([text])[0]

    def ids(self, tokens) -> List[int]:
        if isinstance(tokens, EncodedText):
            return tokens.ids_np            # numpy view: collate copies it without a Python list in between
        unk = self.vocab[self.unk]
        return [self.vocab.get(t, unk) for t in tokens]
