"""CPU: pin the oracle and the host-side code against OUTPUTS OF THE REFERENCE ITSELF.

tests/golden/ref_*.{npz,json} were produced by oracle/make_reference_golden.py, which executes the reference's unmodified
first-party files (MemVul/model_memory.py, custom_PTM_embedder.py, custom_metric.py, predict_memory.py:cal_metrics) in
this container over third-party stand-ins (oracle/ref_shim.py) with the seeded synthetic weights.  These tests read the
fixtures only -- /root/reference is not needed."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import memvul_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF_CASES = ["ref_tiny_same0", "ref_tiny_same1", "ref_base", "ref_tiny_bank130"]


def load_case(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    with open(os.path.join(GOLD, name + ".json")) as f:
        j = json.load(f)
    return z, j, O.BertShape(**j["shape"])


def rows_close(got, want, tol):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g["Issue_Url"] == w["Issue_Url"] and g["label"] == w["label"] and set(g["predict"]) == set(w["predict"])
        for k, v in w["predict"].items():
            assert abs(g["predict"][k] - v) <= tol, (k, g["predict"][k], v)


@pytest.mark.parametrize("name", REF_CASES)
def test_oracle_reproduces_the_reference_run(name):
    z, j, shape = load_case(name)
    sd = O.synthetic_state_dict(shape)
    same = int(z["same_idx"])
    assert same == j["label_vocab"].index("same")
    a_ids, a_mask = torch.from_numpy(z["anchor_ids"]), torch.from_numpy(z["anchor_mask"])
    ids, mask, tids = (torch.from_numpy(z[k]) for k in ("ids", "mask", "type_ids"))
    with torch.no_grad():
        bank = O.build_bank(sd, [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(a_ids.shape[0])], shape)
        ref_bank = torch.from_numpy(z["bank"])
        out = O.memory_forward(sd, ids, mask, tids, ref_bank, same, shape)
    assert float((bank - ref_bank).abs().max()) < 2e-5
    assert float((out["u"] - torch.from_numpy(z["u"])).abs().max()) < 2e-5
    assert float((out["logits"] - torch.from_numpy(z["logits"])).abs().max()) < 2e-5
    assert float((out["p"] - torch.from_numpy(z["p"])).abs().max()) < 2e-6
    # separable form (what the CUDA kernel evaluates) against the reference's cat/Linear form
    sep = O.match_separable(out["u"], ref_bank, sd["_projector.weight"], same)
    assert float((sep["logits"] - torch.from_numpy(z["logits"])).abs().max()) < 2e-5
    rows_close(O.human_readable(out["p"], j["anchor_labels"], j["metadata"], same), j["rows"], 2e-6)


@pytest.mark.parametrize("name", REF_CASES)
def test_host_metrics_and_cal_metrics_match_the_reference(name, tmp_path):
    """custom_metric.py:64-95 + AllenNLP CategoricalAccuracy/FBetaMeasure as ModelMemory.get_metrics reports them
    (model_memory.py:194-217), and predict_memory.py:159-197, both fed with the REFERENCE's probabilities."""
    from memvul_b200.custom_metric import CategoricalAccuracy, FBetaMeasure, SiameseMeasureV1
    from memvul_b200.predict_memory import cal_metrics
    z, j, _ = load_case(name)
    same = int(z["same_idx"])
    p = z["p"].astype(np.float64)
    best = p[np.arange(p.shape[0]), p[:, :, same].argmax(1)]             # probs at the arg-max anchor (model_memory.py:144-147)
    want = j["metrics"]
    acc = CategoricalAccuracy(); acc(best, z["label"])
    assert acc.get_metric(True) == pytest.approx(want["accuracy"])
    for avg, keys in (("weighted", ("precision", "recall", "f1-score")), (None, None)):
        fb = FBetaMeasure(2, average=avg); fb(best, z["label"])
        pr, rc, f1 = fb.get_metric(True).values()
        if avg:
            assert (pr, rc, f1) == pytest.approx(tuple(want[k] for k in keys), abs=1e-6)
        else:
            for i, tok in enumerate(j["label_vocab"]):
                assert (pr[i], rc[i], f1[i]) == pytest.approx((want[f"{tok}_precision"], want[f"{tok}_recall"], want[f"{tok}_f1-score"]), abs=1e-6)
    sm = SiameseMeasureV1(same); sm(best, j["metadata"])
    s = sm.get_metric(True)
    for ours, theirs in (("precision", "s_precision"), ("recall", "s_recall"), ("f1", "s_f1-score"), ("thres", "s_thres"),
                         ("auc", "s_auc"), ("ave_precision_score", "s_ave_precision_score")):
        assert s[ours] == pytest.approx(want[theirs], abs=1e-9), ours
    f = tmp_path / "golden_result.json"
    f.write_text(json.dumps(j["rows"][:2]) + "\n" + json.dumps(j["rows"][2:]) + "\n")
    for thres, ref in j["cal_metrics"].items():
        got = cal_metrics(str(f), thres=float(thres))
        assert set(got) == set(ref)
        for k in ref:
            assert got[k] == pytest.approx(ref[k], abs=1e-12), (thres, k)


def test_fixture_provenance():
    for name in REF_CASES:
        _, j, _ = load_case(name)
        assert "MemVul/model_memory.py" in j["versions"]["reference_files"]


@pytest.mark.parametrize("name", ["ref_single_tiny", "ref_single_c1"])
def test_oracle_single_head_reproduces_the_reference_run(name):
    """MemVul/model_single.py:76-98 executed from /root/reference (config C1 shape: B=4, S=128)."""
    z, j, shape = load_case(name)
    sd = O.synthetic_state_dict(shape, model="single")
    ids, mask, tids = (torch.from_numpy(z[k]) for k in ("ids", "mask", "type_ids"))
    with torch.no_grad():
        out = O.single_forward(sd, ids, mask, tids, shape)
    assert float((out["logits"] - torch.from_numpy(z["logits"])).abs().max()) < 2e-5
    assert float((out["probs"] - torch.from_numpy(z["probs"])).abs().max()) < 2e-6
    loss = torch.nn.functional.cross_entropy(out["logits"], torch.from_numpy(z["label"]))
    assert float(loss) == pytest.approx(float(z["loss"]), abs=1e-6)
    assert [r["predict"] for r in j["rows"]] == [j["label_vocab"][int(k)] for k in out["probs"].argmax(-1)]


def test_host_reader_reproduces_the_reference_reader(tmp_path_factory):
    """MemVul/reader_memory.py executed from /root/reference over the toy files (tests/toy_vocab.py): same instances,
    same order, same word pieces, labels and metadata."""
    from memvul_b200.registrable import DatasetReader, Vocabulary
    from toy_vocab import TOY_VOCAB, write_toy_data
    with open(os.path.join(GOLD, "ref_reader.json")) as f:
        ref = json.load(f)["instances"]
    tmp = tmp_path_factory.mktemp("data")                    # the dispatch is a substring test on the whole path
    vocab_file = os.path.join(str(tmp), "vocab.txt")
    with open(vocab_file, "w") as f:
        f.write("\n".join(TOY_VOCAB) + "\n")
    paths = write_toy_data(tmp)
    reader = DatasetReader.from_params({"type": "reader_memory", "target": "Security_Issue_Full",
                                        "tokenizer": {"type": "pretrained_transformer", "model_name": vocab_file,
                                                      "add_special_tokens": True, "max_length": 16},
                                        "token_indexers": {"tokens": {"type": "pretrained_transformer", "namespace": "tags"}},
                                        "cve_dict_path": paths["cve"]})
    labels = Vocabulary({"labels": ["same", "diff"]})
    reader.index_with(labels)
    for kind in ("golden", "test", "validation"):
        got = list(reader.read(paths[kind]))
        assert len(got) == len(ref[kind])
        for g, w in zip(got, ref[kind]):
            assert list(g["sample1"]["token_ids"]) == w["token_ids"] and list(g["sample1"]["type_ids"]) == w["type_ids"]
            assert g["metadata"] == w["metadata"]
            assert g.get("label_str") == w["label"]
            if w["label"] is not None:
                assert g["label"] == labels.get_token_index(w["label"], "labels")


def test_host_reader_single_reproduces_the_reference_reader(tmp_path_factory):
    """MemVul/reader_single.py executed from /root/reference over the toy files: same instances in the same order (groups
    in first-seen label order, NOT reversed), same word pieces, labels and metadata (BASELINE configs[0] plumbing)."""
    from memvul_b200.registrable import DatasetReader, Vocabulary
    from toy_vocab import TOY_VOCAB, write_toy_data
    with open(os.path.join(GOLD, "ref_reader_single.json")) as f:
        ref = json.load(f)["instances"]
    tmp = tmp_path_factory.mktemp("data")
    vocab_file = os.path.join(str(tmp), "vocab.txt")
    with open(vocab_file, "w") as f:
        f.write("\n".join(TOY_VOCAB) + "\n")
    paths = write_toy_data(tmp)
    reader = DatasetReader.from_params({"type": "reader_single", "target": "Security_Issue_Full",
                                        "tokenizer": {"type": "pretrained_transformer", "model_name": vocab_file,
                                                      "add_special_tokens": True, "max_length": 16},
                                        "token_indexers": {"tokens": {"type": "pretrained_transformer", "namespace": "tags"}}})
    labels = Vocabulary({"class_labels": ["neg", "pos"]})
    reader.index_with(labels)
    for kind in ("test", "validation"):
        got = list(reader.read(paths[kind]))
        assert len(got) == len(ref[kind])
        for g, w in zip(got, ref[kind]):
            assert list(g["sample"]["token_ids"]) == w["token_ids"] and list(g["sample"]["type_ids"]) == w["type_ids"]
            assert g["metadata"] == w["metadata"] and g["label_str"] == w["label"]
            assert g["label"] == labels.get_token_index(w["label"], "class_labels")
    import shutil
    shutil.copy(paths["test"], os.path.join(str(tmp), "train_project.json"))
    with pytest.raises(NotImplementedError):            # the sampled training stream is out of scope
        list(reader.read(os.path.join(str(tmp), "train_project.json")))
