"""The native (C++) batched WordPiece tokenizer against HuggingFace ``tokenizers`` (the backend the reference uses
through AllenNLP, config_memory.json:12-20) and the pure-Python restatement: identical ids on ASCII texts (native path),
identical through the fallback on Unicode / special-token texts, truncation and the [UNK] rules included."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vocab_file(tmp_path_factory):
    rnd = random.Random(7)
    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", ".", ",", "!", "-", "(", ")", "'", "a", "b", "x", "##s", "##ing", "##ed",
             "café", "cafe", "##é"]
    seen = set(words)
    while len(words) < 3000:
        w = "".join(rnd.choice("abcdefghijklmnopqrstuvwxyz0123456789") for _ in range(rnd.randint(1, 7)))
        if rnd.random() < 0.35:
            w = "##" + w
        if w not in seen:
            seen.add(w)
            words.append(w)
    p = tmp_path_factory.mktemp("tok") / "vocab.txt"
    p.write_text("\n".join(words) + "\n", encoding="utf-8")
    return str(p)


def _texts(n=400, seed=3):
    rnd = random.Random(seed)
    alpha = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"
    out = []
    for i in range(n):
        parts = []
        for _ in range(rnd.randint(0, 120)):
            r = rnd.random()
            if r < 0.70:
                parts.append("".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 9))))
            elif r < 0.82:
                parts.append(rnd.choice(".,!-()'\"/\\[]{}<>@#$%^&*+=~`|;:?_"))
            elif r < 0.88:
                parts.append(rnd.choice(["\t", "\n", "\r\n", "  ", "\x0b", "\x00", "\x1f", "\x7f"]))
            elif r < 0.90:
                parts.append("x" * rnd.choice([99, 100, 101, 150]))
            else:
                parts.append(rnd.choice(["don't", "use-after-free", "a.b.c", "[x]", "foo_bar(baz)"]))
        out.append(rnd.choice([" ", "", " "]).join(parts))
    out += ["", " ", "...", "ALLCAPS words Here", "a" * 100, "a" * 101 + " b"]
    return out


def test_native_library_exports_the_header_symbols():
    from memvul_b200 import tokenizer_native as T
    T.build()
    hdr = open(os.path.join(ROOT, "include", "memvul_tok.h")).read()
    declared = sorted(set(re.findall(r"\b(memvul_tok_[a-z0-9_]+)\s*\(", hdr)))
    L = ctypes.CDLL(T.LIB_PATH)
    assert declared == sorted(T.EXPORTS)
    for name in declared:
        assert hasattr(L, name)


@pytest.mark.parametrize("max_length", [None, 16, 64, 512])
def test_native_matches_hf_tokenizers_on_ascii(vocab_file, max_length):
    tokenizers = pytest.importorskip("tokenizers")
    from memvul_b200.tokenizer import FastWordPieceTokenizer, WordPieceTokenizer
    from memvul_b200.tokenizer_native import NativeWordPieceTokenizer
    nat = NativeWordPieceTokenizer(vocab_file, max_length=max_length)
    fast = FastWordPieceTokenizer(vocab_file, max_length=max_length)
    slow = WordPieceTokenizer(vocab_file, max_length=max_length)
    texts = _texts()
    ids, lens = nat.encode_batch(texts)
    assert nat.fallback_count == 0                       # every text took the native path
    for i, (t, enc) in enumerate(zip(texts, fast.tokenize_batch(texts))):
        want = fast.ids(enc)
        got = ids[i, :lens[i]].tolist()
        assert got == want, (t, got, want)
        assert slow.ids(slow.tokenize(t)) == want
        assert int(ids[i, lens[i]:].sum()) == 0           # zero padding
    assert list(nat.tokenize(texts[3])) == list(fast.tokenize(texts[3]))


def test_unicode_and_special_tokens_take_the_fallback_with_identical_results(vocab_file):
    pytest.importorskip("tokenizers")
    from memvul_b200.tokenizer import FastWordPieceTokenizer
    from memvul_b200.tokenizer_native import NativeWordPieceTokenizer
    nat = NativeWordPieceTokenizer(vocab_file, max_length=32)
    fast = FastWordPieceTokenizer(vocab_file, max_length=32)
    texts = ["café heap", "a [MASK] b", "plain ascii", "中文 text", "naïve [SEP] x"]
    ids, lens = nat.encode_batch(texts)
    assert nat.fallback_count == 4
    for i, t in enumerate(texts):
        assert ids[i, :lens[i]].tolist() == fast.ids(fast.tokenize(t)), t


def test_reader_uses_the_native_backend_by_default(vocab_file, tmp_path_factory):
    from memvul_b200.tokenizer import build_tokenizer
    from memvul_b200.tokenizer_native import NativeWordPieceTokenizer
    tok = build_tokenizer({"type": "pretrained_transformer", "model_name": vocab_file, "max_length": 8})
    assert isinstance(tok, NativeWordPieceTokenizer)
    enc = tok.tokenize("Hello, world x")
    assert isinstance(tok.ids(enc), np.ndarray) and len(enc) <= 8
