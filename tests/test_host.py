"""Host-side drop-in surface (no GPU): registered names, constructor keywords, state_dict keys, reader,
tokenizer, collation, metrics, archive loading, output schema.  Contracts cited from the reference."""
import inspect
import json
import os
import tarfile

import numpy as np
import pytest
import torch

import memvul_b200
from memvul_b200 import predict_memory as PM
from memvul_b200.collate import batches, collate_instances
from memvul_b200.custom_metric import SiameseMeasureV1, confusion, find_best_thres
from memvul_b200.registrable import DatasetReader, Metric, Model, TokenEmbedder, Vocabulary
from memvul_b200.synthetic import BERT_TINY, BERT_TINY_H512, build_memory_model, synthetic_state_dict
from memvul_b200.tokenizer import WordPieceTokenizer

from toy_vocab import TOY_VOCAB  # noqa: E402


@pytest.fixture()
def vocab_file(tmp_path):
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(TOY_VOCAB) + "\n", encoding="utf-8")
    return str(p)


def test_registered_names_match_reference():
    """model_memory.py:39, model_single.py:36, reader_memory.py:35, custom_PTM_embedder.py:22, custom_metric.py:55."""
    assert Model.by_name("model_memory") is memvul_b200.ModelMemory
    assert Model.by_name("model_single") is memvul_b200.ModelSingle
    assert DatasetReader.by_name("reader_memory") is memvul_b200.ReaderMemory
    assert TokenEmbedder.by_name("custom_pretrained_transformer") is memvul_b200.PretrainedTransformerEmbedder
    assert Metric.by_name("siamese_measure_v1") is SiameseMeasureV1
    import MemVul                                    # the reference's package name (predict_memory.py:59)
    assert MemVul.ModelMemory is memvul_b200.ModelMemory


def test_constructor_keywords_match_reference():
    want_model = ["vocab", "text_field_embedder", "PTM", "dropout", "label_namespace", "device", "use_header",
                  "temperature", "initializer", "regularizer"]                     # model_memory.py:41-51
    sig = inspect.signature(memvul_b200.ModelMemory.__init__).parameters
    assert list(sig)[1:len(want_model) + 1] == want_model
    # additions are keyword-only and default to the reference's behaviour (header width 512, model_memory.py:70)
    assert all(p.kind is p.KEYWORD_ONLY for p in list(sig.values())[len(want_model) + 1:]) and sig["header_dim"].default == 512
    want_single = ["vocab", "text_field_embedder", "PTM", "dropout", "label_namespace", "device", "initializer", "regularizer"]
    sig = inspect.signature(memvul_b200.ModelSingle.__init__).parameters
    assert list(sig)[1:len(want_single) + 1] == want_single                          # model_single.py:38-46
    assert all(p.kind is p.KEYWORD_ONLY for p in list(sig.values())[len(want_single) + 1:]) and sig["header_dim"].default == 512
    want_emb = ["model_name", "max_length", "sub_module", "train_parameters", "eval_mode", "last_layer_only",
                "override_weights_file", "override_weights_strip_prefix", "gradient_checkpointing", "tokenizer_kwargs",
                "transformer_kwargs", "pretrained_model_path"]                       # custom_PTM_embedder.py:66-81
    got = list(inspect.signature(memvul_b200.PretrainedTransformerEmbedder.__init__).parameters)[1:]
    assert got[:len(want_emb)] == want_emb
    want_reader = ["tokenizer", "same_diff_ratio", "target", "anchor_path", "sample_neg", "train_iter", "token_indexers"]
    assert list(inspect.signature(memvul_b200.ReaderMemory.__init__).parameters)[1:8] == want_reader  # reader_memory.py:38-45
    fwd = list(inspect.signature(memvul_b200.ModelMemory.forward).parameters)[1:]
    assert fwd == ["sample1", "sample2", "label", "metadata"]                     # model_memory.py:118-122
    assert list(inspect.signature(memvul_b200.PretrainedTransformerEmbedder.forward).parameters)[1:5] == \
        ["token_ids", "mask", "type_ids", "segment_concat_mask"]                  # custom_PTM_embedder.py:172-178 (+ kw-only extras)


def test_state_dict_keys_are_the_archive_keys():
    """HF BertModel names under _text_field_embedder.token_embedder_tokens.transformer_model.*, plus
    _bert_pooler.pooler.dense, _projector_single._linear_layers.0, _projector.weight [2,1536]."""
    from memvul_b200.synthetic import BERT_BASE
    model, sd = build_memory_model(BERT_BASE)
    keys = set(model.state_dict().keys())
    assert keys == set(sd.keys())
    p = "_text_field_embedder.token_embedder_tokens.transformer_model."
    for k in [p + "embeddings.word_embeddings.weight", p + "encoder.layer.11.attention.self.query.weight",
              p + "encoder.layer.0.attention.output.LayerNorm.bias", p + "encoder.layer.5.intermediate.dense.weight",
              p + "pooler.dense.weight", "_bert_pooler.pooler.dense.weight", "_projector_single._linear_layers.0.weight",
              "_projector.weight"]:
        assert k in keys
    assert tuple(model.state_dict()["_projector.weight"].shape) == (2, 1536)
    assert tuple(model.state_dict()["_projector_single._linear_layers.0.weight"].shape) == (512, 768)
    assert len(keys) == 5 + 12 * 16 + 2 + 2 + 2 + 1
    single = Model.by_name("model_single")(Vocabulary({"class_labels": ["neg", "pos"]}),
                                          {"token_embedders": {"tokens": {"type": "custom_pretrained_transformer",
                                                                          "model_name": "bert-base-uncased", "pretrained_model_path": ""}}})
    sk = set(single.state_dict().keys())
    assert {"_projector.0._linear_layers.0.weight", "_projector.0._linear_layers.0.bias", "_projector.1.weight"} <= sk


def test_bank_attributes_and_reset_protocol():
    """callbacks.py:48-49 resets the memory by assigning None to both attributes."""
    model, _ = build_memory_model(BERT_TINY)
    assert model._golden_instances_embeddings is None and model._golden_instances_labels is None
    model._golden_instances_embeddings = torch.zeros(3, 64)
    model._golden_instances_labels = ["a", "b", "c"]
    model._golden_instances_embeddings = None
    model._golden_instances_labels = None
    assert model.get_output_dim() == 128 and model.get_output_dim(use_header=True) == 64
    with pytest.raises(NotImplementedError):
        model(sample1=None, sample2=None, label=None, metadata=[{"type": "train"}])


def test_vocabulary_same_idx(tmp_path):
    (tmp_path / "labels.txt").write_text("diff\nsame\n")
    (tmp_path / "non_padded_namespaces.txt").write_text("*labels\n")
    v = Vocabulary.from_files(str(tmp_path))
    assert v.get_token_index("same", "labels") == 1 and v.get_vocab_size("labels") == 2
    assert v.get_index_to_token_vocabulary("labels") == {0: "diff", 1: "same"}


def test_wordpiece_matches_hf_bert_tokenizer(vocab_file):
    """Pinned against HuggingFace's BERT WordPiece (the `tokenizers` backend AllenNLP's tokenizer wraps)."""
    tokenizers = pytest.importorskip("tokenizers")
    hf = tokenizers.BertWordPieceTokenizer(vocab_file, lowercase=True)
    ours = WordPieceTokenizer(vocab_file, add_special_tokens=True, max_length=None)
    texts = ["Buffer overflow in the parser.", "SQL injections, crash when URLTAG is NULL!", "Fixed use-after-free",
             "caf\u00e9 heap   xyz\tunknownword", "a [MASK] b", "", "xyzxyz . !,"]
    for t in texts:
        e = hf.encode(t)
        assert ours.tokenize(t) == e.tokens, t
        assert ours.ids(ours.tokenize(t)) == e.ids
    short = WordPieceTokenizer(vocab_file, max_length=6)
    hf.enable_truncation(max_length=6)
    assert short.tokenize(texts[1]) == hf.encode(texts[1]).tokens


def _write_data(tmp_path, vocab_file):
    from toy_vocab import write_toy_data
    paths = write_toy_data(tmp_path)
    g, c, t, v = paths["golden"], paths["cve"], paths["test"], paths["validation"]
    reader = DatasetReader.from_params({"type": "reader_memory", "target": "Security_Issue_Full",
                                        "tokenizer": {"type": "pretrained_transformer", "model_name": vocab_file,
                                                      "add_special_tokens": True, "max_length": 16},
                                        "token_indexers": {"tokens": {"type": "pretrained_transformer", "namespace": "tags"}},
                                        "cve_dict_path": str(c)})
    return reader, str(g), str(t), str(v)


def test_reader_eval_branches(tmp_path_factory, vocab_file):
    """reader_memory.py:138-162 (dispatch, reversed order => positives first), :231-245 (label, metadata).
    The dispatch is a substring test on the whole path (as in the reference), so the data directory must not
    itself contain "test_" -- pytest's per-test tmp dirs do."""
    tmp_path = tmp_path_factory.mktemp("data")
    reader, g, t, v = _write_data(tmp_path, vocab_file)
    reader.index_with(Vocabulary({"labels": ["same", "diff"]}))
    gold = list(reader.read(g))
    assert [i["metadata"]["instance"][0]["label"] for i in gold] == ["CWE-79", "CWE-120", "CWE-416"]
    assert all(i["metadata"]["type"] == "golden" and i["label"] is None for i in gold)
    assert list(gold[1]["sample1"]["token_ids"]) == [2, 5, 6, 7, 3]          # [CLS] buffer over ##flow [SEP]
    test = list(reader.read(t))
    # dataset = {"neg":[u0,u2], "CWE-120":[u1], "CWE-79":[u3]} (u4 dropped: CWE id None) -> reversed concat
    assert [i["metadata"]["instance"][0]["Issue_Url"] for i in test] == ["u3", "u1", "u2", "u0"]
    assert [i["metadata"]["instance"][0]["label"] for i in test] == ["CWE-79", "CWE-120", "neg", "neg"]
    assert [i["label_str"] for i in test] == ["same", "same", "diff", "diff"] and [i["label"] for i in test] == [0, 0, 1, 1]
    assert all(i["metadata"]["type"] == "unlabel" for i in test)
    assert all(i["metadata"]["type"] == "test" for i in reader.read(v))
    (tmp_path / "train_project.json").write_text((tmp_path / "validation_project.json").read_text())
    with pytest.raises(NotImplementedError):
        list(reader.read(str(tmp_path / "train_project.json")))


def test_collate_pads_to_longest(tmp_path_factory, vocab_file):
    reader, g, t, _ = _write_data(tmp_path_factory.mktemp("data"), vocab_file)
    reader.index_with(Vocabulary({"labels": ["same", "diff"]}))
    inst = list(reader.read(t))
    batch = collate_instances(inst[:3])
    tok = batch["sample1"]["tokens"]
    lens = [len(i["sample1"]["token_ids"]) for i in inst[:3]]
    assert tok["token_ids"].shape == (3, max(lens)) and tok["token_ids"].dtype == torch.int64 and tok["mask"].dtype == torch.bool
    assert tok["mask"].sum(1).tolist() == lens and int(tok["type_ids"].max()) == 0
    assert batch["label"].tolist() == [0, 0, 1] and len(batch["metadata"]) == 3
    assert [len(b) for b in batches(inst, 3)] == [3, 1]


def test_length_bucket_plan_is_a_permutation_with_less_padding():
    from memvul_b200.collate import plan_length_buckets
    g = torch.Generator().manual_seed(0)
    lens = [int([128, 256, 512][i]) for i in torch.randint(0, 3, (1000,), generator=g)]
    plan = plan_length_buckets(lens, 64, window=8)
    assert sorted(i for b in plan for i in b) == list(range(1000)) and all(len(b) <= 64 for b in plan)
    padded = lambda batches: sum(max(lens[i] for i in b) * len(b) for b in batches)
    naive = [list(range(i, min(1000, i + 64))) for i in range(0, 1000, 64)]
    assert padded(plan) < 0.7 * padded(naive)
    assert padded(plan) >= sum(lens)


def test_threshold_sweep_matches_reference_semantics():
    """custom_metric.py:35-52: thresholds np.arange(0.5, 0.9, 0.01); the LAST threshold reaching the best F1 wins."""
    rng = np.random.default_rng(0)
    labels = (rng.random(400) < 0.3).astype(int)
    scores = np.clip(labels * 0.25 + rng.random(400) * 0.75, 0, 1)
    best = find_best_thres(labels, scores)
    ref_best, ref_f1 = None, 0
    for th in np.arange(0.5, 0.9, 0.01):                      # straight restatement of the reference's python loop
        pred = [1 if s >= th else 0 for s in scores]
        tp = sum(1 for p, l in zip(pred, labels) if p == l == 1); fn = sum(1 for p, l in zip(pred, labels) if l == 1 and p != l)
        fp = sum(1 for p, l in zip(pred, labels) if l == 0 and p != l)
        rec = tp / (tp + fn) if tp + fn else 0; pr = tp / (tp + fp) if tp + fp else 0
        f1 = 2 * rec * pr / (rec + pr) if rec + pr else 0
        if f1 >= ref_f1:
            ref_f1, ref_best = f1, th
    assert best["thres"] == pytest.approx(ref_best) and best["f1"] == pytest.approx(ref_f1)
    m = SiameseMeasureV1(same_idx=0)
    meta = [{"instance": [{"label": "neg" if l == 0 else "CWE-1"}]} for l in labels]
    m(np.stack([scores, 1 - scores], 1), meta)
    out = m.get_metric(reset=True)
    assert out["thres"] == pytest.approx(ref_best) and 0.5 < out["auc"] <= 1.0
    assert m.get_metric(reset=True)["f1"] == 0          # cleared
    assert confusion([1, 1, 0, 0], [1, 0, 1, 0]) == {"TP": 1, "FN": 1, "TN": 1, "FP": 1, "precision": 0.5, "recall": 0.5, "f1": 0.5}


def test_human_readable_schema_and_cal_metrics(tmp_path):
    """model_memory.py:169-191 (last anchor of a CWE id wins) -> JSON lines -> predict_memory.py:159-197."""
    model, _ = build_memory_model(BERT_TINY)
    model._golden_instances_labels = ["CWE-1", "CWE-2", "CWE-1"]
    p = [[[0.2, 0.8], [0.7, 0.3], [0.6, 0.4]], [[0.45, 0.55], [0.1, 0.9], [0.3, 0.7]]]
    meta = [{"type": "unlabel", "instance": [{"label": "CWE-2", "Issue_Url": "a"}]},
            {"type": "unlabel", "instance": [{"label": "neg", "Issue_Url": "b"}]}]
    rows = model.make_output_human_readable({"meta": meta, "probs": p})
    assert rows[0]["Issue_Url"] == "a" and rows[0]["label"] == "CWE-2"
    assert rows[0]["predict"] == {"CWE-1": pytest.approx(0.6), "CWE-2": pytest.approx(0.7)}
    assert rows[1]["predict"] == {"CWE-1": pytest.approx(0.3), "CWE-2": pytest.approx(0.1)}
    assert model.make_output_human_readable({"meta": [{"type": "golden"}]}) == {"meta": [{"type": "golden"}]}
    f = tmp_path / "out_result.json"
    f.write_text(json.dumps(rows) + "\n" + json.dumps(rows) + "\n")
    m = PM.cal_metrics(str(f), thres=0.5)
    assert (m["TP"], m["FN"], m["TN"], m["FP"]) == (2, 0, 2, 0) and m["thres"] == 0.5 and m["auc"] == 1.0
    m = PM.cal_metrics(str(f), thres=0.75)
    assert (m["TP"], m["FN"]) == (0, 2)


def test_load_archive_roundtrip(tmp_path, vocab_file):
    """predict_memory.py:62-67: config.json + vocabulary/ + weights.th, with dict overrides."""
    model, sd = build_memory_model(BERT_TINY_H512)             # a config-built model has the reference's 512-wide header
    d = tmp_path / "ser"
    (d / "vocabulary").mkdir(parents=True)
    (d / "vocabulary" / "labels.txt").write_text("diff\nsame\n")
    cfg = {"dataset_reader": {"type": "reader_memory", "tokenizer": {"type": "pretrained_transformer", "model_name": vocab_file, "max_length": 256}},
           "model": {"type": "model_memory", "label_namespace": "labels", "dropout": 0.1, "device": "cuda:0", "use_header": True,
                     "PTM": "bert-base-uncased", "temperature": 0.1,
                     "text_field_embedder": {"token_embedders": {"tokens": {
                         "type": "custom_pretrained_transformer", "model_name": "bert-base-uncased", "train_parameters": True,
                         "pretrained_model_path": "further_pretrain/out_wwm/",
                         "transformer_kwargs": {"vocab_size": 1024, "hidden_size": 128, "num_hidden_layers": 2,
                                                "num_attention_heads": 2, "intermediate_size": 512}}}}},
           "validation_data_loader": {"batch_size": 512, "shuffle": False}}
    (d / "config.json").write_text(json.dumps(cfg))
    state = dict(sd)
    state["_text_field_embedder.token_embedder_tokens.transformer_model.embeddings.position_ids"] = torch.arange(512)[None]
    torch.save(state, d / "weights.th")
    tar = tmp_path / "model.tar.gz"
    with tarfile.open(tar, "w:gz") as t:
        for n in ("config.json", "weights.th", "vocabulary"):
            t.add(d / n, arcname=n)
    over = {"validation_dataset_reader": {"type": "reader_memory", "tokenizer": {"type": "pretrained_transformer",
                                                                                  "model_name": vocab_file, "max_length": 512}},
            "model": {"device": "cpu"}}
    arc = PM.load_archive(str(tar), overrides=over)
    assert arc.model._same_idx == 1 and arc.config["model"]["device"] == "cpu" and arc.config["model"]["temperature"] == 0.1
    for k, v in arc.model.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert arc.dataset_reader._tokenizer.max_length == 256 and arc.validation_dataset_reader._tokenizer.max_length == 512


def test_custom_validation_callback_refreshes_the_bank(tmp_path_factory, vocab_file):
    """callbacks.py:43-53: on_epoch empties the memory, then re-encodes anchors[:128] and anchors[128:] through
    forward_on_instances; registered under the reference's names."""
    from memvul_b200.registrable import TrainerCallback
    assert TrainerCallback.by_name("custom_validation").__name__ == "CustomValidation"
    assert TrainerCallback.by_name("reset_dataloader").__name__ == "ResetLoader"
    tmp_path = tmp_path_factory.mktemp("data")
    reader, g, _, _ = _write_data(tmp_path, vocab_file)
    reader.index_with(Vocabulary({"labels": ["same", "diff"]}))
    many = {f"CWE-{i}": "buffer overflow" for i in range(130)}
    (tmp_path / "big_golden_anchors.json").write_text(json.dumps(many))

    class FakeModel:
        def __init__(self):
            self._golden_instances_embeddings, self._golden_instances_labels = "stale", ["stale"]
            self.calls, self.evaled = [], False
        def eval(self):
            self.evaled = True
        def forward_on_instances(self, inst):
            assert self._golden_instances_labels is None or self._golden_instances_labels != ["stale"]
            self.calls.append(len(inst))
            self._golden_instances_labels = (self._golden_instances_labels or []) + [i["metadata"]["instance"][0]["label"] for i in inst]

    class FakeTrainer:
        pass
    tr = FakeTrainer(); tr.model = FakeModel(); tr.data_loader = FakeTrainer(); tr.data_loader._instances = [1, 2]
    cb = TrainerCallback.from_params({"type": "custom_validation", "anchor_path": str(tmp_path / "big_golden_anchors.json")},
                                     data_reader=reader)
    cb.on_epoch(tr, {}, 0, True)
    assert tr.model.evaled and tr.model.calls == [128, 2] and len(tr.model._golden_instances_labels) == 130
    assert tr.model._golden_instances_embeddings is None          # the fake never sets it: proves the reset happened
    cb.on_epoch(tr, {}, 1, True)
    assert tr.model.calls == [128, 2, 128, 2] and len(tr.model._golden_instances_labels) == 130
    TrainerCallback.by_name("reset_dataloader")().on_epoch(tr, {}, 0, True)
    assert tr.data_loader._instances is None
    with pytest.raises(ValueError):
        TrainerCallback.by_name("custom_validation")(anchor_path=g)
