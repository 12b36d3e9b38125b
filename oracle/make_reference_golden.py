"""TEST INFRASTRUCTURE -- golden vectors produced by EXECUTING the reference's own files in this container.

    python oracle/make_reference_golden.py            # needs /root/reference; writes tests/golden/ref_*.npz / .json

The reference (panshengyi/MemVul) cannot be imported as is: AllenNLP 2.4.0, `overrides`, matplotlib and spaCy are not in
the image.  `oracle/ref_shim.py` supplies stand-ins for those THIRD-PARTY packages only; the first-party
files are loaded unmodified from /root/reference:
    MemVul/model_memory.py        ModelMemory.__init__ / forward_gold_instances / forward (test branch :133-147) /
                                  make_output_human_readable (:169-191) / get_metrics (:194-217)
    MemVul/custom_PTM_embedder.py PretrainedTransformerEmbedder.__init__ / forward (:172-242) over transformers.BertModel
    MemVul/custom_metric.py       SiameseMeasureV1, find_best_thres, cal_f1
    predict_memory.py             cal_metrics (:159-197), model_measure (:117-156)
    MemVul/model_single.py        ModelSingle.__init__ / forward (:76-98) / make_output_human_readable / get_metrics
    MemVul/reader_memory.py       ReaderMemory.__init__ / read_dataset / _read eval branches / text_to_instance
The weights are the seeded synthetic state_dict (memvul_b200/synthetic.py, seed 2021) loaded through the reference
model's own `load_state_dict`, i.e. under the reference's parameter names; inputs are the seeded synthetic ids.  What
is stored per case: inputs, the anchor bank, the header output u, the projector logits (forward hooks), `output_dict
["probs"]`, the human-readable rows, `get_metrics(reset=True)`, and cal_metrics' output file.  The fixtures travel; this
script and /root/reference do not need to exist on the GPU box.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

# name -> (shape kwargs, report lengths, S, anchor lengths, labels-namespace order, report labels, report CWE labels)
CASES = {
    "ref_tiny_same0": (dict(vocab_size=1024, hidden=128, layers=2, heads=2, intermediate=512, max_pos=512, header=512),
                       [40, 7, 22, 2, 33, 40], 40, [9, 33, 64, 17, 40, 5, 64], ["same", "diff"]),
    "ref_tiny_same1": (dict(vocab_size=1024, hidden=128, layers=2, heads=2, intermediate=512, max_pos=512, header=512),
                       [64, 64, 3, 50, 31], 64, [12, 64, 30], ["diff", "same"]),
    "ref_base": (dict(), [48, 20, 5, 37], 48, [16, 48, 9, 30, 25], ["same", "diff"]),
    # 130 anchors: the memory is built as 128 + 2 (predict_memory.py:81-83), longest anchor differs per chunk
    "ref_tiny_bank130": (dict(vocab_size=1024, hidden=128, layers=2, heads=2, intermediate=512, max_pos=512, header=512),
                         [30, 11, 24], 30, [5 + (7 * i) % 28 for i in range(128)] + [40, 6], ["same", "diff"]),
}


# MemVul-m (model_single.py), config C1 shape B=4, S=128: name -> (shape kwargs, lengths, S, class_labels order)
SINGLE_CASES = {
    "ref_single_tiny": (dict(vocab_size=1024, hidden=128, layers=2, heads=2, intermediate=512, max_pos=512, header=512),
                        [128, 9, 64, 100], 128, ["neg", "pos"]),
    "ref_single_c1": (dict(), [128, 9, 64, 100], 128, ["pos", "neg"]),
}


def _tokens(ids, mask, tids):
    return {"tokens": {"token_ids": ids, "mask": mask, "type_ids": tids}}


def run_case(name: str) -> None:
    from memvul_b200.synthetic import BertShape, synthetic_ids, synthetic_state_dict
    import transformers
    from oracle import ref_shim
    kw, lens, S, alens, label_vocab = CASES[name]
    shape = BertShape(**kw)
    ref_shim.install(hidden=shape.hidden, vocab_size=shape.vocab_size)
    mm, emb_mod, met_mod, drv = ref_shim.import_reference(REFERENCE)
    sd = synthetic_state_dict(shape)

    with tempfile.TemporaryDirectory() as tmp:
        cfg = transformers.BertConfig(vocab_size=shape.vocab_size, hidden_size=shape.hidden, num_hidden_layers=shape.layers,
                                      num_attention_heads=shape.heads, intermediate_size=shape.intermediate,
                                      max_position_embeddings=shape.max_pos, type_vocab_size=shape.type_vocab,
                                      layer_norm_eps=shape.ln_eps, hidden_act="gelu")
        transformers.BertModel(cfg).save_pretrained(os.path.join(tmp, "out_wwm"))       # random: overwritten below
        embedder = emb_mod.PretrainedTransformerEmbedder(model_name="bert-base-uncased", train_parameters=True,
                                                        pretrained_model_path=os.path.join(tmp, "out_wwm"))
    vocab = ref_shim.Vocabulary({"labels": label_vocab})
    tfe = ref_shim.BasicTextFieldEmbedder({"tokens": embedder})
    model = mm.ModelMemory(vocab, tfe, dropout=0.1, device="cpu", use_header=True)
    own = model.state_dict()
    unexpected = sorted(set(sd) - set(own))
    missing = sorted(k for k in set(own) - set(sd) if not k.endswith("position_ids"))
    assert not unexpected and not missing, (unexpected, missing)         # the archive key names of SURVEY 8b hold
    model.load_state_dict(sd, strict=False)
    model.eval()                                                         # predict_memory.py:75

    cap = {}
    model._projector.register_forward_hook(lambda m, i, o: cap.__setitem__("logits", o.detach().clone()))
    model._projector_single.register_forward_hook(lambda m, i, o: cap.__setitem__("u", o.detach().clone()))

    G, B = len(alens), len(lens)
    a_ids, a_mask, a_tids = synthetic_ids(G, max(alens), lens=alens, seed=31, vocab_size=shape.vocab_size)
    ids, mask, tids = synthetic_ids(B, S, lens=lens, seed=32, vocab_size=shape.vocab_size)
    cwe = [f"CWE-{(i * 7) % 5}" for i in range(G)] if name != "ref_tiny_bank130" else [f"CWE-{i}" for i in range(G)]
    rep_label = ["neg" if i % 3 else cwe[i % G] for i in range(B)]
    label = torch.tensor([vocab.get_token_index("diff" if l == "neg" else "same", "labels") for l in rep_label])
    meta = [{"type": "unlabel", "instance": [{"label": rep_label[i], "Issue_Url": f"https://example.test/issue/{i}"}]}
            for i in range(B)]
    with torch.no_grad():
        for c0 in range(0, G, 128):                                      # predict_memory.py:81-83
            c1 = min(G, c0 + 128) if c0 == 0 else G
            Sg = int(a_mask[c0:c1].sum(1).max())                         # AllenNLP collate pads to the chunk's longest
            out = model(sample1=_tokens(a_ids[c0:c1, :Sg].contiguous(), a_mask[c0:c1, :Sg].contiguous(),
                                        a_tids[c0:c1, :Sg].contiguous()),
                        metadata=[{"type": "golden", "instance": [{"label": cwe[g]}]} for g in range(c0, c1)])
            assert out == {}
            if c1 == G:
                break
        bank = model._golden_instances_embeddings.clone()
        out = model(sample1=_tokens(ids, mask, tids), label=label, metadata=meta)
    p = np.asarray(out["probs"], dtype=np.float64)                      # list[B][G][2] python floats (model_memory.py:143)
    rows = model.make_output_human_readable(out)
    metrics = model.get_metrics(reset=True)

    # predict_memory.py:159-197 on the predictions file evaluate() would have written (one JSON array per batch per line)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "test_results"))
        with open(os.path.join(tmp, "test_results", "golden_result.json"), "w") as f:
            f.write(json.dumps(rows[:2]) + "\n" + json.dumps(rows[2:]) + "\n")
        drv.DATA_PATH = tmp
        voted = {}
        for thres in (0.5, 0.62):
            drv.cal_metrics("golden_result", thres=thres)
            voted[str(thres)] = json.load(open(os.path.join(tmp, "test_results", "golden_metric_all.json")))

    np.savez_compressed(os.path.join(GOLD, name + ".npz"), ids=ids.numpy(), mask=mask.numpy(), type_ids=tids.numpy(),
                        anchor_ids=a_ids.numpy(), anchor_mask=a_mask.numpy(), label=label.numpy(),
                        bank=bank.numpy(), u=cap["u"].numpy(), logits=cap["logits"].numpy(), p=p.astype(np.float32),
                        same_idx=np.int64(model._same_idx))
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump({"shape": kw, "label_vocab": label_vocab, "anchor_labels": model._golden_instances_labels,
                   "metadata": meta, "rows": rows, "metrics": metrics, "cal_metrics": voted,
                   "versions": {"torch": torch.__version__, "transformers": transformers.__version__,
                                "reference_files": ["MemVul/model_memory.py", "MemVul/custom_PTM_embedder.py",
                                                    "MemVul/custom_metric.py", "predict_memory.py"]}},
                  f, indent=1, default=float)
    print(f"{name}: bank {tuple(bank.shape)} p {p.shape} same_idx {model._same_idx} metrics {json.dumps(metrics, default=float)[:120]}")


def run_single_case(name: str) -> None:
    """MemVul/model_single.py: ModelSingle.forward (:76-98), make_output_human_readable (:100-110), get_metrics."""
    import importlib
    from memvul_b200.synthetic import BertShape, synthetic_ids, synthetic_state_dict
    import transformers
    from oracle import ref_shim
    kw, lens, S, label_vocab = SINGLE_CASES[name]
    shape = BertShape(**kw)
    ref_shim.install(hidden=shape.hidden, vocab_size=shape.vocab_size)
    _, emb_mod, _, _ = ref_shim.import_reference(REFERENCE)
    ms = importlib.import_module("MemVul.model_single")
    sd = synthetic_state_dict(shape, model="single")
    with tempfile.TemporaryDirectory() as tmp:
        cfg = transformers.BertConfig(vocab_size=shape.vocab_size, hidden_size=shape.hidden, num_hidden_layers=shape.layers,
                                      num_attention_heads=shape.heads, intermediate_size=shape.intermediate,
                                      max_position_embeddings=shape.max_pos, type_vocab_size=shape.type_vocab,
                                      layer_norm_eps=shape.ln_eps, hidden_act="gelu")
        transformers.BertModel(cfg).save_pretrained(os.path.join(tmp, "out_wwm"))
        embedder = emb_mod.PretrainedTransformerEmbedder(model_name="bert-base-uncased",
                                                        pretrained_model_path=os.path.join(tmp, "out_wwm"))
    vocab = ref_shim.Vocabulary({"class_labels": label_vocab})
    model = ms.ModelSingle(vocab, ref_shim.BasicTextFieldEmbedder({"tokens": embedder}), dropout=0.1, device="cpu")
    own = model.state_dict()
    unexpected = sorted(set(sd) - set(own))
    missing = sorted(k for k in set(own) - set(sd) if not k.endswith("position_ids"))
    assert not unexpected and not missing, (unexpected, missing)
    model.load_state_dict(sd, strict=False)
    model.eval()
    cap = {}
    model._projector.register_forward_hook(lambda m, i, o: cap.__setitem__("logits", o.detach().clone()))
    B = len(lens)
    ids, mask, tids = synthetic_ids(B, S, lens=lens, vocab_size=shape.vocab_size)
    rep = ["neg", "pos", "pos", "neg"]
    label = torch.tensor([vocab.get_token_index(l, "class_labels") for l in rep])
    meta = [{"instance": {"Issue_Url": f"https://example.test/issue/{i}", "label": rep[i]}} for i in range(B)]
    with torch.no_grad():
        out = model(_tokens(ids, mask, tids), label=label, metadata=meta)
    rows = model.make_output_human_readable(out)
    metrics = model.get_metrics(reset=True)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), ids=ids.numpy(), mask=mask.numpy(), type_ids=tids.numpy(),
                        label=label.numpy(), logits=cap["logits"].numpy(), probs=np.asarray(out["probs"], dtype=np.float32),
                        loss=np.float32(out["loss"]))
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump({"shape": kw, "label_vocab": label_vocab, "metadata": meta, "rows": rows, "metrics": metrics,
                   "versions": {"torch": torch.__version__, "transformers": transformers.__version__,
                                "reference_files": ["MemVul/model_single.py", "MemVul/custom_PTM_embedder.py"]}},
                  f, indent=1, default=float)
    print(f"{name}: probs {np.asarray(out['probs']).round(4).tolist()} loss {float(out['loss']):.6f}")


def run_reader_case(name: str = "ref_reader") -> None:
    """MemVul/reader_memory.py: __init__ (:38-71), read_dataset (:73-113), _read eval branches (:138-162),
    text_to_instance (:195-246), over the toy vocabulary / data files the host tests use (tests/toy_vocab.py)."""
    import importlib
    from oracle import ref_shim
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from toy_vocab import TOY_VOCAB, write_toy_data
    ref_shim.install()
    ref_shim.import_reference(REFERENCE)
    rm = importlib.import_module("MemVul.reader_memory")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        vocab_file = os.path.join(tmp, "vocab.txt")
        with open(vocab_file, "w") as f:
            f.write("\n".join(TOY_VOCAB) + "\n")
        paths = write_toy_data(tmp)                                   # golden / test / validation / CVE dict
        cwd = os.getcwd()
        os.chdir(tmp)                                                 # the reference opens "xxx" + "CVE_dict.json" (:64-66)
        try:
            os.replace(paths["cve"], os.path.join(tmp, "xxxCVE_dict.json"))
            tok = ref_shim.PretrainedTransformerTokenizer(vocab_file, add_special_tokens=True, max_length=16)
            reader = rm.ReaderMemory(tokenizer=tok, target="Security_Issue_Full", anchor_path=paths["golden"], sample_neg=0.1,
                                     token_indexers={"tokens": ref_shim.PretrainedTransformerIndexer(vocab_file, namespace="tags")})
            for kind in ("golden", "test", "validation"):
                rows = []
                for inst in reader.read(paths[kind]):
                    f = inst.fields
                    idx = f["sample1"]._token_indexers["tokens"].tokens_to_indices(f["sample1"].tokens)
                    rows.append({"tokens": [t.text for t in f["sample1"].tokens], "token_ids": idx["token_ids"],
                                 "type_ids": idx["type_ids"], "label": f["label"].label if "label" in f else None,
                                 "metadata": f["metadata"].metadata})
                out[kind] = rows
        finally:
            os.chdir(cwd)
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump({"instances": out, "reference_files": ["MemVul/reader_memory.py", "MemVul/util.py"]}, f, indent=1)
    print(name, {k: len(v) for k, v in out.items()})


def run_reader_single_case(name: str = "ref_reader_single") -> None:
    """MemVul/reader_single.py: __init__ (:33-51), read_dataset (:53-71), _read evaluation branches (:73-94),
    text_to_instance (:112-126), over the same toy vocabulary / data files as ``ref_reader``."""
    import importlib
    from oracle import ref_shim
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from toy_vocab import TOY_VOCAB, write_toy_data
    ref_shim.install()
    ref_shim.import_reference(REFERENCE)
    rs = importlib.import_module("MemVul.reader_single")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        vocab_file = os.path.join(tmp, "vocab.txt")
        with open(vocab_file, "w") as f:
            f.write("\n".join(TOY_VOCAB) + "\n")
        paths = write_toy_data(tmp)
        tok = ref_shim.PretrainedTransformerTokenizer(vocab_file, add_special_tokens=True, max_length=16)
        reader = rs.ReaderSingle(tokenizer=tok, target="Security_Issue_Full", sample_neg=0.1,
                                 token_indexers={"tokens": ref_shim.PretrainedTransformerIndexer(vocab_file, namespace="tags")})
        for kind in ("test", "validation"):
            rows = []
            for inst in reader.read(paths[kind]):
                f = inst.fields
                idx = f["sample"]._token_indexers["tokens"].tokens_to_indices(f["sample"].tokens)
                rows.append({"tokens": [t.text for t in f["sample"].tokens], "token_ids": idx["token_ids"],
                             "type_ids": idx["type_ids"], "label": f["label"].label, "metadata": f["metadata"].metadata})
            out[kind] = rows
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump({"instances": out, "reference_files": ["MemVul/reader_single.py"]}, f, indent=1)
    print(name, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    import subprocess
    if len(sys.argv) > 1 and sys.argv[1] == "ref_reader_single":
        run_reader_single_case()
    elif len(sys.argv) > 1 and sys.argv[1] == "ref_reader":
        run_reader_case()
    elif len(sys.argv) > 1:
        (run_single_case if sys.argv[1] in SINGLE_CASES else run_case)(sys.argv[1])
    else:                           # one process per case: ref_shim.SETTINGS and the reference modules are per-process
        for name in list(CASES) + list(SINGLE_CASES) + ["ref_reader", "ref_reader_single"]:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name])
