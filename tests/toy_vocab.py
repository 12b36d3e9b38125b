"""Toy WordPiece vocabulary shared by the reader / tokenizer / driver tests."""
TOY_VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "buffer", "over", "##flow", "in", "the", "parser", ".", ",",
             "sql", "injection", "##s", "crash", "when", "url", "##tag", "is", "null", "a", "b", "fix", "##ed", "!", "use",
             "after", "free", "-", "heap", "cafe", "x", "##y", "##z"]
