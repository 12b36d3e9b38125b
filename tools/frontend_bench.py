"""Front-end throughput: text -> JSON lines through ``predict_memory.test_siamese`` against the
pre-tokenised end-to-end rate, on a synthetic file of N issue reports (~512 word pieces each) and 129 anchors.

    python tools/frontend_bench.py [--n 10000] [--out profiles/r02_frontend.json] [--cpu-only]

Legs:  tokenise   reader.read() alone (lazy chunks, HF `tokenizers` Rust backend unless MEMVUL_TOKENIZER=python)
       text_e2e   test_siamese: JSON file -> reader -> prefetch thread -> GPU -> JSON lines (wall clock, reports/s)
       pretok_e2e evaluate() over the already tokenised instances (the same model, memory and output file)
"""
import argparse
import json
import os
import random
import sys
import tarfile
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_vocab(path, n=30522, seed=1):
    rnd = random.Random(seed)
    words = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    seen = set(words)
    letters = "abcdefghijklmnopqrstuvwxyz"
    while len(words) < n:
        w = "".join(rnd.choice(letters) for _ in range(rnd.randint(3, 9)))
        if rnd.random() < 0.2:
            w = "##" + w
        if w not in seen:
            seen.add(w)
            words.append(w)
    with open(path, "w") as f:
        f.write("\n".join(words) + "\n")
    return [w for w in words[104:] if not w.startswith("##")]


def make_data(d, words, n, seed=2):
    rnd = random.Random(seed)
    rows = []
    for i in range(n):
        k = rnd.choice((120, 250, 510)) if os.environ.get("FE_MIXED") else 510
        body = " ".join(rnd.choice(words) for _ in range(k))
        rows.append({"Issue_Url": f"u{i}", "Issue_Title": rnd.choice(words), "Issue_Body": body,
                     "Security_Issue_Full": int(i % 301 == 0), "CVE_ID": f"CVE-{i % 50}"})
    anchors = {f"CWE-{i}": " ".join(rnd.choice(words) for _ in range(rnd.randint(20, 200))) for i in range(129)}
    cve = {f"CVE-{i}": {"CWE_ID": f"CWE-{i % 129}"} for i in range(50)}
    paths = {"test": os.path.join(d, "test_project.json"), "golden": os.path.join(d, "CWE_anchor_golden_project.json"),
             "cve": os.path.join(d, "CVE_dict.json")}
    for k, obj in (("test", rows), ("golden", anchors), ("cve", cve)):
        with open(paths[k], "w") as f:
            json.dump(obj, f)
    return paths


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--out", default="")
    ap.add_argument("--cpu-only", action="store_true", help="tokenisation leg only (no GPU needed)")
    args = ap.parse_args()
    from memvul_b200.registrable import DatasetReader
    d = tempfile.mkdtemp(prefix="fe_data_")
    vocab_file = os.path.join(d, "vocab.txt")
    words = make_vocab(vocab_file)
    paths = make_data(d, words, args.n)
    tok = {"type": "pretrained_transformer", "model_name": vocab_file, "add_special_tokens": True, "max_length": 512}
    res = {"reports": args.n, "host_cpus": os.cpu_count(), "tokenizer_backend": os.environ.get("MEMVUL_TOKENIZER", "tokenizers (Rust)")}
    reader = DatasetReader.from_params({"type": "reader_memory", "tokenizer": dict(tok), "cve_dict_path": paths["cve"]})
    t0 = time.perf_counter()
    n_tok = n_inst = 0
    for inst in reader.read(paths["test"]):
        n_inst += 1
        n_tok += len(inst["sample1"]["token_ids"])
    dt = time.perf_counter() - t0
    res["tokenise"] = {"reports_per_s": n_inst / dt, "tokens_per_s": n_tok / dt, "seconds": dt, "mean_tokens": n_tok / n_inst}
    print("tokenise:", res["tokenise"], flush=True)
    if not args.cpu_only:
        from memvul_b200 import predict_memory as PM
        from memvul_b200.synthetic import BERT_BASE, synthetic_state_dict
        (os.makedirs(os.path.join(d, "vocabulary")))
        with open(os.path.join(d, "vocabulary", "labels.txt"), "w") as f:
            f.write("same\ndiff\n")
        cfg = {"dataset_reader": {"type": "reader_memory", "tokenizer": tok, "cve_dict_path": paths["cve"]},
               "validation_dataset_reader": {"type": "reader_memory", "tokenizer": tok},
               "model": {"type": "model_memory", "device": "cuda:0", "text_field_embedder": {"token_embedders": {"tokens": {
                   "type": "custom_pretrained_transformer", "model_name": "bert-base-uncased", "pretrained_model_path": ""}}}},
               "validation_data_loader": {"batch_size": 64, "shuffle": False}}
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump(cfg, f)
        torch.save(synthetic_state_dict(BERT_BASE), os.path.join(d, "weights.th"))
        out_file = os.path.join(d, "out_result.json")
        # text -> JSON lines (archive given as the extracted directory: tar + gzip of 438 MB of weights is not the subject)
        PM.test_siamese(d, paths["test"], paths["golden"], predictions_output_file=out_file, batch_size=64, cuda_device=0)   # warm-up (JIT-free, but cold caches)
        t0 = time.perf_counter()
        PM.test_siamese(d, paths["test"], paths["golden"], predictions_output_file=out_file, batch_size=64, cuda_device=0)
        dt_all = time.perf_counter() - t0
        # the same with the model already loaded and the memory built: reader -> GPU -> JSON only
        arc = PM.load_archive(d, cuda_device=0)
        model = arc.model.eval()
        for r in (arc.dataset_reader, arc.validation_dataset_reader):
            r.index_with(model.vocab)
        PM.build_memory(model, list(arc.validation_dataset_reader.read(paths["golden"])))
        dev = torch.device("cuda:0")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        arc.dataset_reader._dataset.clear()
        PM.evaluate(model, arc.dataset_reader.read(paths["test"]), 64, dev, predictions_output_file=out_file)
        torch.cuda.synchronize()
        dt_text = time.perf_counter() - t0
        insts = list(arc.dataset_reader.read(paths["test"]))       # tokenised now (cached in the reader)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        PM.evaluate(model, insts, 64, dev, predictions_output_file=out_file)
        torch.cuda.synchronize()
        dt_pre = time.perf_counter() - t0
        lines = sum(1 for _ in open(out_file))
        res["text_e2e"] = {"reports_per_s": n_inst / dt_text, "seconds": dt_text,
                           "what": "reader (lazy tokenisation) -> prefetch thread -> GPU -> JSON lines, model and memory already built"}
        res["pretok_e2e"] = {"reports_per_s": n_inst / dt_pre, "seconds": dt_pre, "what": "same, instances already tokenised"}
        res["test_siamese_total"] = {"reports_per_s": n_inst / dt_all, "seconds": dt_all,
                                     "what": "whole driver call incl. archive load (438 MB weights), weight packing, memory build"}
        res["text_over_pretok"] = dt_pre / dt_text
        res["output_lines"] = lines
        print(json.dumps(res), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
