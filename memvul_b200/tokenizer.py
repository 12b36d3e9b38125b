"""BERT WordPiece tokenisation for the reader (AllenNLP's
``PretrainedTransformerTokenizer`` / ``PretrainedTransformerIndexer`` pair in config_memory.json:12-27).

``WordPieceTokenizer`` is a self-contained uncased BERT tokenizer (basic clean-up / lower-casing /
accent stripping / punctuation splitting, then greedy longest-match WordPiece) that needs only a
``vocab.txt``; tests pin it against ``transformers.BertTokenizer`` on the same vocabulary.
``add_special_tokens`` wraps in [CLS] ... [SEP] and ``max_length`` truncates the TOTAL length, as the
AllenNLP tokenizer does with HF ``truncation=True``.
"""
from __future__ import annotations

import os
import unicodedata
from typing import Dict, List, Optional


def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0x2A700 <= cp <= 0x2B73F
            or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


class WordPieceTokenizer:
    def __init__(self, vocab_file: str, lowercase: bool = True, add_special_tokens: bool = True,
                 max_length: Optional[int] = None, unk: str = "[UNK]", cls: str = "[CLS]", sep: str = "[SEP]",
                 max_chars_per_word: int = 100) -> None:
        with open(vocab_file, encoding="utf-8") as f:
            toks = [ln.rstrip("\n") for ln in f]
        self.vocab: Dict[str, int] = {t: i for i, t in enumerate(toks) if t != ""}
        self.lowercase, self.add_special, self.max_length = lowercase, add_special_tokens, max_length
        self.unk, self.cls, self.sep, self.max_chars = unk, cls, sep, max_chars_per_word
        self.never_split = {"[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]"}

    # --- basic tokenizer ---
    def _clean(self, text: str) -> str:
        out = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or (unicodedata.category(ch) in ("Cc", "Cf") and ch not in "\t\n\r"):
                continue
            if ch in " \t\n\r" or unicodedata.category(ch) == "Zs":
                out.append(" ")
            elif _is_cjk(cp):
                out.append(f" {ch} ")
            else:
                out.append(ch)
        return "".join(out)

    def _basic(self, text: str) -> List[str]:
        text = unicodedata.normalize("NFC", self._clean(text))
        words: List[str] = []
        for tok in text.split():
            if tok in self.never_split:
                words.append(tok)
                continue
            if self.lowercase:
                tok = tok.lower()
                tok = "".join(c for c in unicodedata.normalize("NFD", tok) if unicodedata.category(c) != "Mn")
            cur = ""
            for ch in tok:
                if _is_punct(ch):
                    if cur:
                        words.append(cur)
                        cur = ""
                    words.append(ch)
                else:
                    cur += ch
            if cur:
                words.append(cur)
        return words

    def _wordpiece(self, word: str) -> List[str]:
        if len(word) > self.max_chars:
            return [self.unk]
        pieces, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end]
                if start > 0:
                    sub = "##" + sub
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [self.unk]
            pieces.append(cur)
            start = end
        return pieces

    def tokenize(self, text: str) -> List[str]:
        pieces: List[str] = []
        for w in self._basic(text):
            pieces.extend([w] if w in self.never_split else self._wordpiece(w))
        if self.add_special:
            if self.max_length is not None:
                pieces = pieces[: max(self.max_length - 2, 0)]
            return [self.cls] + pieces + [self.sep]
        return pieces if self.max_length is None else pieces[: self.max_length]

    def tokenize_batch(self, texts: List[str]) -> List[List[str]]:
        return [self.tokenize(t) for t in texts]

    def ids(self, tokens: List[str]) -> List[int]:
        unk = self.vocab[self.unk]
        return [self.vocab.get(t, unk) for t in tokens]


class TokenList(list):
    """Word pieces of one text that remember their ids (the Rust backend returns both at once)."""
    ids: List[int]


class FastWordPieceTokenizer:
    """The same tokenisation through HuggingFace ``tokenizers`` (Rust): what the reference itself runs -- AllenNLP's
    ``PretrainedTransformerTokenizer`` wraps the HF *fast* BERT tokenizer (config_memory.json:12-20,
    MemVul/reader_memory.py:88).  ``tokenize_batch`` encodes a list of texts on all host cores with the GIL released,
    which is what lets the reader keep up with the GPU (the pure-Python ``WordPieceTokenizer`` does ~1.5 k reports/s
    on one core; the encoder consumes ~9 k/s per GPU).  Pinned against ``WordPieceTokenizer`` and HF's own
    ``BertWordPieceTokenizer`` in tests/test_host.py."""

    def __init__(self, vocab_file: str, lowercase: bool = True, add_special_tokens: bool = True,
                 max_length: Optional[int] = None, unk: str = "[UNK]", cls: str = "[CLS]", sep: str = "[SEP]",
                 max_chars_per_word: int = 100) -> None:
        from tokenizers import Tokenizer, models, normalizers, pre_tokenizers, processors
        tok = Tokenizer(models.WordPiece.from_file(vocab_file, unk_token=unk, max_input_chars_per_word=max_chars_per_word))
        tok.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None,
                                                    lowercase=lowercase)
        tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
        special = [t for t in ("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]") if tok.token_to_id(t) is not None]
        tok.add_special_tokens(special)                       # never split / lower-cased, like BERT's never_split
        if add_special_tokens:
            tok.post_processor = processors.TemplateProcessing(
                single=f"{cls} $A {sep}", special_tokens=[(cls, tok.token_to_id(cls)), (sep, tok.token_to_id(sep))])
        if max_length is not None:
            tok.enable_truncation(max_length=max_length)      # total length, special tokens included (HF truncation=True)
        self._tok = tok
        self.lowercase, self.add_special, self.max_length = lowercase, add_special_tokens, max_length
        self.unk, self.cls, self.sep = unk, cls, sep
        self.vocab = tok.get_vocab()

    @staticmethod
    def _wrap(enc) -> TokenList:
        t = TokenList(enc.tokens)
        t.ids = list(enc.ids)
        return t

    def tokenize(self, text: str) -> TokenList:
        return self._wrap(self._tok.encode(text))

    def tokenize_batch(self, texts: List[str]) -> List[TokenList]:
        return [self._wrap(e) for e in self._tok.encode_batch(list(texts))]

    def ids(self, tokens: List[str]) -> List[int]:
        if isinstance(tokens, TokenList) and hasattr(tokens, "ids"):
            return tokens.ids
        unk = self.vocab[self.unk]
        return [self.vocab.get(t, unk) for t in tokens]


def build_tokenizer(spec):
    """``spec``: an object with ``tokenize``/``ids`` or the config block
    ``{"type": "pretrained_transformer", "model_name": ..., "add_special_tokens": true, "max_length": 256}``.
    ``model_name`` may be a directory containing ``vocab.txt`` or the file itself; the bare hub name
    ("bert-base-uncased") resolves through ``$MEMVUL_VOCAB`` because this image is offline.
    Backend (``MEMVUL_TOKENIZER`` = native | tokenizers | python, default native): the in-tree C++ batched tokenizer
    (memvul_b200/tokenizer_native.py, falls back per text to a Unicode-complete backend), HF ``tokenizers`` (Rust --
    the reference's own backend), or the pure-Python ``WordPieceTokenizer``.  All three agree token for token."""
    if hasattr(spec, "tokenize"):
        return spec
    spec = dict(spec or {})
    name = spec.get("model_name", "bert-base-uncased")
    cands = [name, os.path.join(name, "vocab.txt"), os.environ.get("MEMVUL_VOCAB", "")]
    for c in cands:
        if c and os.path.isfile(c):
            kw = dict(add_special_tokens=spec.get("add_special_tokens", True), max_length=spec.get("max_length"))
            backend = os.environ.get("MEMVUL_TOKENIZER", "native")
            if backend == "native":
                try:
                    from .tokenizer_native import NativeWordPieceTokenizer
                    return NativeWordPieceTokenizer(c, **kw)
                except (OSError, RuntimeError):     # no compiler / library: the Rust or Python backend still works
                    pass
            if backend != "python":
                try:
                    return FastWordPieceTokenizer(c, **kw)
                except ImportError:
                    pass
            return WordPieceTokenizer(c, **kw)
    raise FileNotFoundError(f"no vocab.txt for tokenizer {name!r}: pass a directory/file or set MEMVUL_VOCAB "
                            "(pretrained vocabularies cannot be downloaded in this environment)")
