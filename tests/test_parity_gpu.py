"""End-to-end GPU parity of the drop-in (ModelMemory / ModelSingle / embedder, through the C ABI) against the
CPU oracle and the committed golden vectors.  Gates (BASELINE.json north_star):
    |logits - oracle| <= 1e-3 absolute;  arg-max anchor and pos/neg label identical."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3          # north_star: "match logits within 1e-3 absolute"


def _dev(ids, mask, tids=None):
    d = {"token_ids": ids.cuda(), "mask": mask.cuda()}
    d["type_ids"] = (torch.zeros_like(ids) if tids is None else tids).cuda()
    return {"tokens": d}


def _build_bank(model, a_ids, a_mask, chunk=128):
    n = a_ids.shape[0]
    for c0 in range(0, n, chunk):
        ids, mask = a_ids[c0:c0 + chunk], a_mask[c0:c0 + chunk]
        S = int(mask.sum(1).max())
        model.forward_gold_instances(_dev(ids[:, :S].contiguous(), mask[:, :S].contiguous()),
                                     [{"type": "golden", "instance": [{"label": f"CWE-{c0 + i}"}]} for i in range(ids.shape[0])])


def _meta(n, kind="unlabel"):
    return [{"type": kind, "instance": [{"label": "neg" if i % 3 else f"CWE-{i}", "Issue_Url": f"url/{i}"}]} for i in range(n)]


@pytest.mark.parametrize("name", ["tiny_ragged", "tiny_same1", "base_small"])
def test_model_memory_matches_golden(name):
    from memvul_b200.synthetic import build_memory_model
    from oracle.make_golden import CASES
    shape, lens, S, alens, same = CASES[name]
    z = np.load(os.path.join(GOLD, name + ".npz"))
    model, _ = build_memory_model(shape, same_first=(same == 0), device="cuda")
    assert model._same_idx == same
    with torch.no_grad():
        _build_bank(model, torch.from_numpy(z["anchor_ids"]), torch.from_numpy(z["anchor_mask"]))
        out = model(sample1=_dev(torch.from_numpy(z["ids"]), torch.from_numpy(z["mask"])),
                    label=torch.zeros(len(lens), dtype=torch.int64, device="cuda"), metadata=_meta(len(lens)))
    dev = out["native"]["device"]
    assert float((model._golden_instances_embeddings.cpu() - torch.from_numpy(z["bank"])).abs().max()) < TOL
    assert float((dev["u"].cpu() - torch.from_numpy(z["u"])).abs().max()) < TOL
    assert float((dev["logits"].cpu() - torch.from_numpy(z["logits"])).abs().max()) < TOL
    p = np.asarray(out["probs"].tolist(), dtype=np.float32)
    assert p.shape == z["p"].shape and np.abs(p - z["p"]).max() < TOL
    # gates on EVERY row (memvul_b200/parity.py): labels identical outside the tolerance band, the chosen anchor is a
    # maximiser of the golden probabilities up to the observed error, identical where the golden top-2 gap is clear
    from memvul_b200.parity import gate_report
    g = gate_report(dev["logits"].cpu().numpy(), p, out["native"]["best_idx"].tolist(), z["logits"], z["p"], same,
                    thresholds=(0.5, 0.55), tol=TOL)
    print(name, {k: g[k] for k in ("max_logit_err", "min_margin", "rows_excluded", "argmax_clear_rows", "argmax_mismatch_all")})
    assert g["ok"] and g["label_mismatch_outside_tol"] == 0 and g["argmax_not_maximiser"] == 0, g
    rows = model.make_output_human_readable(out)
    assert len(rows) == len(lens) and set(rows[0]) == {"Issue_Url", "label", "predict"} and len(rows[0]["predict"]) == len(alens)
    json.dumps(rows)


def test_bert_base_batch_against_oracle_with_labels():
    """bert-base, 8 ragged issue reports x 129 anchors (the real memory size), oracle computed on CPU in seconds."""
    from memvul_b200.synthetic import BERT_BASE, build_memory_model, synthetic_ids
    from oracle import memvul_oracle as O
    model, sd = build_memory_model(BERT_BASE, device="cuda")
    g = torch.Generator().manual_seed(5)
    alens = torch.randint(8, 65, (129,), generator=g).tolist()
    a_ids, a_mask, _ = synthetic_ids(129, 64, lens=alens, seed=21)
    lens = [128, 100, 64, 17, 128, 90, 2, 77]
    ids, mask, tids = synthetic_ids(8, 128, lens=lens, seed=22)
    with torch.no_grad():
        _build_bank(model, a_ids, a_mask)                      # 128 + 1 chunks, as predict_memory.py:81-83
        out = model(sample1=_dev(ids, mask, tids), label=torch.ones(8, dtype=torch.int64, device="cuda"), metadata=_meta(8))
        bank = O.build_bank(sd, [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(129)])
        ref = O.memory_forward(sd, ids, mask, tids, bank, model._same_idx)
    dev = out["native"]["device"]
    err = float((dev["logits"].cpu() - ref["logits"]).abs().max())
    assert err < TOL, err
    from memvul_b200.parity import gate_report
    g = gate_report(dev["logits"].cpu().numpy(), np.asarray(out["probs"].tolist()), out["native"]["best_idx"].tolist(),
                    ref["logits"].numpy(), ref["p"].numpy(), model._same_idx, thresholds=(0.5,), tol=TOL)
    print(f"logits max err {err:.2e}; gates {g}")
    assert g["ok"] and g["labels"]["0.5"]["mismatch_rows"] == 0 and g["argmax_clear_rows"] >= 1, g   # labels: every row
    m = model.get_metrics(reset=True)
    assert 0.0 <= m["accuracy"] <= 1.0 and "s_thres" in m and "same_f1-score" in m


def test_full_size_properties_c2():
    """BASELINE configs[1] size (bert-base, S=512, B=64, G=129): size-independent properties."""
    from memvul_b200.synthetic import BERT_BASE, build_memory_model, synthetic_ids
    from oracle import memvul_oracle as O
    model, sd = build_memory_model(BERT_BASE, device="cuda")
    g = torch.Generator().manual_seed(9)
    alens = torch.randint(16, 129, (129,), generator=g).tolist()
    a_ids, a_mask, _ = synthetic_ids(129, 128, lens=alens, seed=31)
    lens = torch.randint(300, 513, (64,), generator=g).tolist()
    lens[0], lens[1] = 512, 512
    ids, mask, tids = synthetic_ids(64, 512, lens=lens, seed=32)
    ids[1], mask[1] = ids[0], mask[0]                               # duplicate rows -> identical outputs
    with torch.no_grad():
        _build_bank(model, a_ids, a_mask)
        r1 = model.match_batch(_dev(ids, mask, tids))
        l1, p1, b1 = r1["logits"].clone(), r1["probs"].clone(), r1["best_idx"].clone()
        perm = torch.randperm(64, generator=g)
        r2 = model.match_batch(_dev(ids[perm].contiguous(), mask[perm].contiguous(), tids))
        assert torch.equal(r2["logits"], l1[perm.cuda()]), "batch order must not change any sample's result"
        assert torch.equal(l1[0], l1[1]) and int(b1[0]) == int(b1[1])
        assert torch.equal(b1.long(), p1[:, :, model._same_idx].argmax(1))
        assert float((p1.sum(-1) - 1).abs().max()) < 1e-6
        # shrinking the padding of a short batch does not change its rows
        short = [i for i in range(64) if lens[i] <= 384][:8]
        if short:
            S2 = max(lens[i] for i in short)
            ra = model.match_batch(_dev(ids[short].contiguous(), mask[short].contiguous()))["logits"].clone()
            rb = model.match_batch(_dev(ids[short][:, :S2].contiguous(), mask[short][:, :S2].contiguous()))["logits"]
            assert float((ra - rb).abs().max()) < 1e-5
    # (all 64 rows are compared with the oracle in tests/test_configs_gpu.py::test_c2_all_64_rows_against_the_oracle)


def test_embedder_interface_and_errors():
    from memvul_b200 import native
    from memvul_b200.synthetic import BERT_TINY, build_memory_model, synthetic_ids
    from oracle import memvul_oracle as O
    model, sd = build_memory_model(BERT_TINY, device="cuda")
    emb = model._text_field_embedder
    ids, mask, tids = synthetic_ids(3, 33, lens=[33, 4, 20], vocab_size=1024)
    tids[:, 10:] = 1
    tids = tids * mask
    with torch.no_grad():
        hid = emb(_dev(ids, mask, tids))
        ref = O.embedder_forward(sd, ids, mask, tids, O.BERT_TINY)
    assert hid.shape == (3, 33, 128) and emb.get_output_dim() == 128
    assert float((hid.cpu() - ref)[mask].abs().max()) < 5e-3
    bad = mask.clone(); bad[1, 20] = True
    with torch.no_grad(), pytest.raises(ValueError):
        model.forward_gold_instances(_dev(ids, bad), _meta(3, "golden"))
    with pytest.raises(RuntimeError):
        model(sample1=_dev(ids, mask), metadata=_meta(3))               # empty memory
    with pytest.raises(ValueError):
        emb.embedder("tokens")(ids.cuda(), mask.cuda()[:, :5])
    # weights changed in place -> packed fp16 copy is rebuilt
    with torch.no_grad():
        h0 = emb(_dev(ids, mask)).clone()
        model.state_dict()[O.EMB + "embeddings.LayerNorm.bias"].add_(0.5)
        assert not torch.allclose(emb(_dev(ids, mask)), h0)


@pytest.mark.parametrize("shape_name,B,S,lens", [("tiny", 5, 200, [200, 130, 7, 64, 129]), ("base", 3, 256, [256, 100, 31])])
def test_cls_only_last_layer_equals_full(shape_name, B, S, lens):
    """MEMVUL_ENC_CLS_ONLY skips rows the path never reads; the [CLS] rows must equal the full forward's."""
    from memvul_b200 import native
    from memvul_b200.synthetic import BERT_BASE, BERT_TINY, EMB, synthetic_ids, synthetic_state_dict
    shape = BERT_TINY if shape_name == "tiny" else BERT_BASE
    w = native.PackedBert(synthetic_state_dict(shape), EMB, torch.device("cuda"))
    ids, mask, _ = synthetic_ids(B, S, lens=lens, vocab_size=shape.vocab_size)
    lens_t, _ = native.mask_to_lens(mask.cuda())
    full = native.encoder_forward(w, ids.cuda(), lens_t)
    cls = native.encoder_forward(w, ids.cuda(), lens_t, cls_only=True)
    assert float((full[:, 0] - cls[:, 0]).abs().max()) < 1e-5


def test_model_single_matches_oracle():
    from memvul_b200.custom_PTM_embedder import PretrainedTransformerEmbedder
    from memvul_b200.model_single import ModelSingle
    from memvul_b200.modules import BasicTextFieldEmbedder
    from memvul_b200.registrable import Vocabulary
    from memvul_b200.synthetic import BERT_TINY, config_lite, load_into, synthetic_ids, synthetic_state_dict
    from oracle import memvul_oracle as O
    sd = synthetic_state_dict(BERT_TINY, model="single")
    emb = PretrainedTransformerEmbedder("bert-base-uncased", pretrained_model_path="", config=config_lite(BERT_TINY))
    model = ModelSingle(Vocabulary({"class_labels": ["neg", "pos"]}), BasicTextFieldEmbedder({"tokens": emb}), header_dim=BERT_TINY.header)
    load_into(model, sd)
    model.eval().cuda()
    ids, mask, tids = synthetic_ids(4, 128, lens=[128, 9, 64, 100], vocab_size=1024)   # config C1 shape: B=4, S=128
    label = torch.tensor([0, 1, 1, 0])
    with torch.no_grad():
        out = model(_dev(ids, mask, tids), label=label.cuda(),
                    metadata=[{"instance": {"Issue_Url": f"u{i}", "label": "neg"}} for i in range(4)])
        ref = O.single_forward(sd, ids, mask, tids, O.BERT_TINY)
    assert float((torch.tensor(out["probs"]) - ref["probs"]).abs().max()) < TOL
    assert float(out["loss"]) == pytest.approx(float(torch.nn.functional.cross_entropy(ref["logits"], label)), abs=TOL)
    rows = model.make_output_human_readable(out)
    assert rows[0].keys() == {"Issue_Url", "label", "predict", "prob"}


def test_predict_driver_end_to_end(tmp_path_factory):
    """predict_memory.test_siamese flow on a toy archive: archive -> bank (128+rest) -> batches -> JSON lines -> cal_metrics."""
    import tarfile
    from memvul_b200 import predict_memory as PM
    from memvul_b200.synthetic import BERT_TINY_H512, synthetic_state_dict
    from toy_vocab import TOY_VOCAB
    d = tmp_path_factory.mktemp("arch")
    vocab_file = d / "vocab.txt"
    vocab_file.write_text("\n".join(TOY_VOCAB) + "\n")
    (d / "vocabulary").mkdir()
    (d / "vocabulary" / "labels.txt").write_text("same\ndiff\n")
    tok = {"type": "pretrained_transformer", "model_name": str(vocab_file), "add_special_tokens": True, "max_length": 32}
    cve = d / "CVE_dict.json"
    cve.write_text(json.dumps({f"CVE-{i}": {"CWE_ID": f"CWE-{i % 3}"} for i in range(20)}))
    cfg = {"dataset_reader": {"type": "reader_memory", "tokenizer": tok, "cve_dict_path": str(cve)},
           "validation_dataset_reader": {"type": "reader_memory", "tokenizer": tok},
           "model": {"type": "model_memory", "device": "cuda:0", "text_field_embedder": {"token_embedders": {"tokens": {
               "type": "custom_pretrained_transformer", "model_name": "bert-base-uncased", "pretrained_model_path": "",
               "transformer_kwargs": {"vocab_size": 1024, "hidden_size": 128, "num_hidden_layers": 2,
                                      "num_attention_heads": 2, "intermediate_size": 512}}}}},
           "validation_data_loader": {"batch_size": 4, "shuffle": False}}
    (d / "config.json").write_text(json.dumps(cfg))
    torch.save(synthetic_state_dict(BERT_TINY_H512), d / "weights.th")
    with tarfile.open(d / "model.tar.gz", "w:gz") as t:
        for n in ("config.json", "weights.th", "vocabulary"):
            t.add(d / n, arcname=n)
    words = ["buffer", "overflow", "parser", "sql", "injection", "crash", "null", "heap", "free", "fix"]
    (d / "CWE_anchor_golden_project.json").write_text(json.dumps({f"CWE-{i}": " ".join(words[i:i + 3]) for i in range(3)}))
    rows = [{"Issue_Url": f"u{i}", "Issue_Title": words[i % 10], "Issue_Body": " ".join(words[(i * 3) % 7:(i * 3) % 7 + 4]),
             "Security_Issue_Full": int(i % 4 == 0), "CVE_ID": f"CVE-{i}"} for i in range(11)]
    (d / "test_project.json").write_text(json.dumps(rows))
    res = d / "out_result.json"
    metrics = PM.test_siamese(str(d / "model.tar.gz"), str(d / "test_project.json"), str(d / "CWE_anchor_golden_project.json"),
                              predictions_output_file=str(res), batch_size=4, cuda_device=0)
    lines = [json.loads(l) for l in res.read_text().splitlines()]
    assert [len(l) for l in lines] == [4, 4, 3]
    assert [r["Issue_Url"] for r in lines[0]][:3] == ["u8", "u4", "u0"]          # positives first (reversed groups)
    assert all(set(r["predict"]) == {"CWE-0", "CWE-1", "CWE-2"} for l in lines for r in l)
    assert "s_f1-score" in metrics and "accuracy" in metrics
    m = PM.cal_metrics(str(res), thres=0.5)
    assert m["TP"] + m["FN"] == 3 and m["TN"] + m["FP"] == 8
    # length-bucketed batching writes the same file (rows are restored to data order)
    res2 = d / "out_result_bucketed.json"
    PM.test_siamese(str(d / "model.tar.gz"), str(d / "test_project.json"), str(d / "CWE_anchor_golden_project.json"),
                    predictions_output_file=str(res2), batch_size=4, cuda_device=0, bucket_by_length=True)
    lines2 = [json.loads(l) for l in res2.read_text().splitlines()]
    assert [[r["Issue_Url"] for r in l] for l in lines2] == [[r["Issue_Url"] for r in l] for l in lines]
    for la, lb in zip(lines, lines2):
        for ra, rb in zip(la, lb):
            assert all(abs(ra["predict"][k] - rb["predict"][k]) < 1e-5 for k in ra["predict"])


def test_in_training_validation_entry_after_weight_update():
    """custom_trainer.py:506-618 runs the `type == "test"` branch after callbacks.py:43-53 has rebuilt
    the memory with the CURRENT weights.  An in-place parameter update (what an optimiser step does) must reach the
    fp16 packed copy; bank, logits and the validation metrics are checked against the oracle on the updated weights."""
    from memvul_b200.callbacks import CustomValidation
    from memvul_b200.custom_metric import SiameseMeasureV1
    from memvul_b200.synthetic import BERT_TINY, build_memory_model, synthetic_ids
    from oracle import memvul_oracle as O
    shape = BERT_TINY
    model, _ = build_memory_model(shape, device="cuda")
    alens = [9, 33, 64, 17, 40, 5, 64]
    a_ids, a_mask, _ = synthetic_ids(len(alens), 64, lens=alens, seed=31, vocab_size=shape.vocab_size)
    anchors = [{"sample1": {"token_ids": a_ids[i][a_mask[i]].tolist(), "type_ids": [0] * alens[i]}, "label": None,
                "metadata": {"type": "golden", "instance": [{"label": f"CWE-{i}"}]}} for i in range(len(alens))]

    class Reader:
        def read(self, path):
            return iter(anchors)

    class Trainer:
        pass
    tr = Trainer(); tr.model = model
    cb = CustomValidation("golden_anchors.json", data_reader=Reader())
    cb.on_epoch(tr, {}, 0, True)
    bank0 = model._golden_instances_embeddings.clone()
    g = torch.Generator(device="cuda").manual_seed(3)
    with torch.no_grad():                                        # "optimiser step": every parameter moves in place
        for p in model.parameters():
            p.add_(torch.randn(p.shape, device=p.device, generator=g) * 0.02 * p.abs().mean().clamp_min(1e-3))
    cb.on_epoch(tr, {}, 1, True)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref_bank = O.build_bank(sd, [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(len(alens))], shape)
    assert model._golden_instances_labels == [f"CWE-{i}" for i in range(len(alens))]
    assert float((model._golden_instances_embeddings.cpu() - ref_bank).abs().max()) < TOL
    assert float((model._golden_instances_embeddings - bank0).abs().max()) > 1e-4          # the refresh is not a no-op
    lens = [64, 20, 3, 50, 64, 31]
    ids, mask, tids = synthetic_ids(len(lens), 64, lens=lens, seed=32, vocab_size=shape.vocab_size)
    label = torch.tensor([0, 1, 1, 0, 1, 1])                     # index into the labels namespace
    with torch.no_grad():
        out = model(sample1=_dev(ids, mask, tids), label=label.cuda(), metadata=_meta(len(lens), kind="test"))
    ref = O.memory_forward(sd, ids, mask, tids, ref_bank, model._same_idx, shape)
    assert float((out["native"]["device"]["logits"].cpu() - ref["logits"]).abs().max()) < TOL
    got = model.get_metrics(reset=True)
    # custom_metric.py:64-72 feeds (gold = label != "neg", probs[b][same_idx]); accuracy is over the arg-max anchor's [B,2]
    meta = _meta(len(lens), kind="test")
    want = SiameseMeasureV1(model._same_idx)
    want(ref["probs"], meta)
    w = want.get_metric(reset=True)
    ps_got = torch.tensor(out["native"]["best_probs"].tolist())
    assert float((ps_got - ref["probs"]).abs().max()) < TOL
    margin = float((ref["probs"][:, 0] - ref["probs"][:, 1]).abs().min())
    if margin > 2 * TOL:                                          # no sample sits on the decision boundary
        assert got["accuracy"] == pytest.approx(float((ref["probs"].argmax(-1) == label).float().mean()))
        assert got["s_thres"] == pytest.approx(w["thres"]) and got["s_f1-score"] == pytest.approx(w["f1"])
        assert got["s_auc"] == pytest.approx(w["auc"], abs=1e-6)
    print(f"validation entry: decision margin {margin:.3e}, accuracy {got['accuracy']:.3f}, s_thres {got['s_thres']}")


@pytest.mark.parametrize("name", ["ref_tiny_same0", "ref_tiny_same1", "ref_base", "ref_tiny_bank130"])
def test_cuda_path_matches_the_reference_run(name, tmp_path):
    """The drop-in against outputs of the REFERENCE's own files executed in the build container
    (oracle/make_reference_golden.py): bank, u, logits, probs, human-readable rows, get_metrics, cal_metrics."""
    from memvul_b200.predict_memory import cal_metrics
    from memvul_b200.synthetic import BertShape, build_memory_model
    z = np.load(os.path.join(GOLD, name + ".npz"))
    with open(os.path.join(GOLD, name + ".json")) as f:
        j = json.load(f)
    same = int(z["same_idx"])
    model, _ = build_memory_model(BertShape(**j["shape"]), same_first=(j["label_vocab"][0] == "same"), device="cuda")
    assert model._same_idx == same
    a_ids, a_mask = torch.from_numpy(z["anchor_ids"]), torch.from_numpy(z["anchor_mask"])
    G = a_ids.shape[0]
    with torch.no_grad():
        for c0, c1 in ((0, min(G, 128)), (128, G)):                     # predict_memory.py:81-83
            if c0 >= c1:
                continue
            S = int(a_mask[c0:c1].sum(1).max())
            assert model(sample1=_dev(a_ids[c0:c1, :S].contiguous(), a_mask[c0:c1, :S].contiguous()),
                         metadata=[{"type": "golden", "instance": [{"label": l}]} for l in j["anchor_labels"][c0:c1]]) == {}
        out = model(sample1=_dev(torch.from_numpy(z["ids"]), torch.from_numpy(z["mask"]), torch.from_numpy(z["type_ids"])),
                    label=torch.from_numpy(z["label"]).cuda(), metadata=j["metadata"])
    assert model._golden_instances_labels == j["anchor_labels"]
    dev = out["native"]["device"]
    errs = {"bank": float((model._golden_instances_embeddings.cpu() - torch.from_numpy(z["bank"])).abs().max()),
            "u": float((dev["u"].cpu() - torch.from_numpy(z["u"])).abs().max()),
            "logits": float((dev["logits"].cpu() - torch.from_numpy(z["logits"])).abs().max())}
    p = np.asarray(out["probs"].tolist(), dtype=np.float32)
    errs["p"] = float(np.abs(p - z["p"]).max())
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
    ps = z["p"][:, :, same]
    from memvul_b200.parity import gate_report
    gr = gate_report(dev["logits"].cpu().numpy(), p, out["native"]["best_idx"].tolist(), z["logits"], z["p"], same,
                     thresholds=(0.5,), tol=TOL)
    assert gr["ok"] and gr["argmax_not_maximiser"] == 0 and gr["label_mismatch_outside_tol"] == 0, gr
    rows = model.make_output_human_readable(out)
    for g, w in zip(rows, j["rows"]):
        assert g["Issue_Url"] == w["Issue_Url"] and g["label"] == w["label"] and set(g["predict"]) == set(w["predict"])
        assert max(abs(g["predict"][k] - w["predict"][k]) for k in w["predict"]) < TOL
    best_ref = z["p"][np.arange(len(ps)), ps.argmax(1)]
    if np.abs(best_ref[:, 0] - best_ref[:, 1]).min() > 2 * TOL:         # nobody on the arg-max boundary: exact metrics
        m = model.get_metrics(reset=True)
        for k in ("accuracy", "precision", "recall", "f1-score"):
            assert m[k] == pytest.approx(j["metrics"][k], abs=1e-6), k
        sweep = np.arange(0.5, 0.9, 0.01)                                # custom_metric.py:35-52 thresholds
        if np.abs(best_ref[:, same][:, None] - sweep[None]).min() > TOL:  # no score within TOL of a sweep threshold
            for k in ("s_precision", "s_recall", "s_f1-score", "s_thres"):
                assert m[k] == pytest.approx(j["metrics"][k], abs=1e-6), k
        order_gap = np.diff(np.sort(best_ref[:, same])).min() if len(best_ref) > 1 else 1.0
        if order_gap > 2 * TOL:                                           # ranking metrics depend on the order only
            assert m["s_auc"] == pytest.approx(j["metrics"]["s_auc"], abs=1e-9)
    f = tmp_path / "golden_result.json"
    f.write_text(json.dumps(rows[:2]) + "\n" + json.dumps(rows[2:]) + "\n")
    for thres, ref in j["cal_metrics"].items():
        vote_ref = np.max(ps, axis=1)
        if np.abs(vote_ref - float(thres)).min() > TOL:
            got = cal_metrics(str(f), thres=float(thres))
            for k in ("TP", "FN", "TN", "FP", "f1"):
                assert got[k] == pytest.approx(ref[k]), (thres, k)


@pytest.mark.parametrize("name", ["ref_single_tiny", "ref_single_c1"])
def test_model_single_matches_the_reference_run(name):
    """MemVul-m (config C1: B=4, S=128) against MemVul/model_single.py executed in the build container."""
    from memvul_b200.custom_PTM_embedder import PretrainedTransformerEmbedder
    from memvul_b200.model_single import ModelSingle
    from memvul_b200.modules import BasicTextFieldEmbedder
    from memvul_b200.registrable import Vocabulary
    from memvul_b200.synthetic import BertShape, config_lite, load_into, synthetic_state_dict
    z = np.load(os.path.join(GOLD, name + ".npz"))
    with open(os.path.join(GOLD, name + ".json")) as f:
        j = json.load(f)
    shape = BertShape(**j["shape"])
    emb = PretrainedTransformerEmbedder("bert-base-uncased", pretrained_model_path="", config=config_lite(shape))
    model = ModelSingle(Vocabulary({"class_labels": j["label_vocab"]}), BasicTextFieldEmbedder({"tokens": emb}), header_dim=shape.header)
    load_into(model, synthetic_state_dict(shape, model="single"))
    model.eval().cuda()
    with torch.no_grad():
        out = model(_dev(torch.from_numpy(z["ids"]), torch.from_numpy(z["mask"]), torch.from_numpy(z["type_ids"])),
                    label=torch.from_numpy(z["label"]).cuda(), metadata=j["metadata"])
    err = float((out["logits_device"].cpu() - torch.from_numpy(z["logits"])).abs().max())
    assert err < TOL and float(np.abs(np.asarray(out["probs"]) - z["probs"]).max()) < TOL
    assert float(out["loss"]) == pytest.approx(float(z["loss"]), abs=TOL)
    rows = model.make_output_human_readable(out)
    margin = float(np.abs(z["probs"][:, 0] - z["probs"][:, 1]).min())
    for g, w in zip(rows, j["rows"]):
        assert g["Issue_Url"] == w["Issue_Url"] and g["label"] == w["label"] and abs(g["prob"] - w["prob"]) < TOL
        if margin > 2 * TOL:
            assert g["predict"] == w["predict"]
    if margin > 2 * TOL:
        m = model.get_metrics(reset=True)
        for k, v in j["metrics"].items():
            assert m[k] == pytest.approx(v, abs=1e-6), k
    print(f"{name}: logits err {err:.2e}, decision margin {margin:.2e}")


def test_use_header_false_matches_the_oracle():
    """model_memory.py:69,101: without the header the 768-wide pooled vectors are matched directly
    (``_projector`` = Linear(3*768 -> 2)); unused by the shipped configs, but part of the constructor contract."""
    from memvul_b200.custom_PTM_embedder import PretrainedTransformerEmbedder
    from memvul_b200.model_memory import ModelMemory
    from memvul_b200.modules import BasicTextFieldEmbedder
    from memvul_b200.parity import gate_report
    from memvul_b200.registrable import Vocabulary
    from memvul_b200.synthetic import BERT_TINY, config_lite, synthetic_ids, synthetic_state_dict
    from oracle import memvul_oracle as O
    shape = BERT_TINY
    sd = {k: v for k, v in synthetic_state_dict(shape).items() if not k.startswith("_projector")}
    g = torch.Generator().manual_seed(11)
    sd["_projector.weight"] = torch.randn(2, 3 * shape.hidden, generator=g) * 0.03
    emb = PretrainedTransformerEmbedder("bert-base-uncased", pretrained_model_path="", config=config_lite(shape))
    model = ModelMemory(Vocabulary({"labels": ["same", "diff"]}), BasicTextFieldEmbedder({"tokens": emb}), use_header=False)
    model.load_state_dict(sd)
    model.eval().cuda()
    assert model.get_output_dim() == shape.hidden and not hasattr(model, "_projector_single")
    alens = [9, 33, 64, 17, 40, 5, 64]
    a_ids, a_mask, _ = synthetic_ids(len(alens), 64, lens=alens, seed=31, vocab_size=shape.vocab_size)
    lens = [64, 20, 3, 50, 64, 31]
    ids, mask, tids = synthetic_ids(len(lens), 64, lens=lens, seed=32, vocab_size=shape.vocab_size)
    with torch.no_grad():
        model.forward_gold_instances(_dev(a_ids, a_mask), [{"type": "golden", "instance": [{"label": f"CWE-{i}"}]} for i in range(len(alens))])
        out = model(sample1=_dev(ids, mask, tids), metadata=_meta(len(lens)))
        bank = O.instance_forward(sd, a_ids, a_mask, None, shape, use_header=False)
        u = O.instance_forward(sd, ids, mask, tids, shape, use_header=False)
        ref = O.match(u, bank, sd["_projector.weight"], model._same_idx)
    assert tuple(model._golden_instances_embeddings.shape) == (len(alens), shape.hidden)
    dev = out["native"]["device"]
    assert float((model._golden_instances_embeddings.cpu() - bank).abs().max()) < TOL
    gr = gate_report(dev["logits"].cpu().numpy(), np.asarray(out["probs"].tolist()), out["native"]["best_idx"].tolist(),
                     ref["logits"].numpy(), ref["p"].numpy(), model._same_idx, thresholds=(0.5,), tol=TOL)
    assert gr["ok"] and gr["argmax_not_maximiser"] == 0, gr
    rows = model.make_output_human_readable(out)
    assert len(rows) == len(lens) and len(rows[0]["predict"]) == len(alens)


def test_predict_single_driver_end_to_end(tmp_path_factory):
    """BASELINE configs[0] plumbing (predict_single.py:46-97 + reader_single): archive -> reader -> batches of 4 ->
    ModelSingle -> JSON lines -> cal_metrics, checked against the oracle's probabilities row by row."""
    import tarfile
    from memvul_b200 import predict_single as PS
    from memvul_b200.synthetic import BERT_TINY_H512, synthetic_state_dict
    from memvul_b200.tokenizer import build_tokenizer
    from oracle import memvul_oracle as O
    from toy_vocab import TOY_VOCAB
    d = tmp_path_factory.mktemp("arch1")
    vocab_file = d / "vocab.txt"
    vocab_file.write_text("\n".join(TOY_VOCAB) + "\n")
    (d / "vocabulary").mkdir()
    (d / "vocabulary" / "class_labels.txt").write_text("neg\npos\n")
    tok = {"type": "pretrained_transformer", "model_name": str(vocab_file), "add_special_tokens": True, "max_length": 128}
    cfg = {"dataset_reader": {"type": "reader_single", "tokenizer": tok},
           "model": {"type": "model_single", "device": "cuda:0", "text_field_embedder": {"token_embedders": {"tokens": {
               "type": "custom_pretrained_transformer", "model_name": "bert-base-uncased", "pretrained_model_path": "",
               "transformer_kwargs": {"vocab_size": 1024, "hidden_size": 128, "num_hidden_layers": 2,
                                      "num_attention_heads": 2, "intermediate_size": 512}}}}},
           "validation_data_loader": {"batch_size": 4, "shuffle": False}}
    (d / "config.json").write_text(json.dumps(cfg))
    sd = synthetic_state_dict(BERT_TINY_H512, model="single")
    torch.save(sd, d / "weights.th")
    with tarfile.open(d / "model.tar.gz", "w:gz") as t:
        for n in ("config.json", "weights.th", "vocabulary"):
            t.add(d / n, arcname=n)
    words = ["buffer", "overflow", "parser", "sql", "injection", "crash", "null", "heap", "free", "fix"]
    rows = [{"Issue_Url": f"u{i}", "Issue_Title": words[i % 10], "Issue_Body": " ".join(words[(i * 3) % 7:(i * 3) % 7 + 4]),
             "Security_Issue_Full": int(i % 4 == 0)} for i in range(10)]
    (d / "test_project.json").write_text(json.dumps(rows))
    res = d / "single_result.json"
    metrics = PS.test(str(d / "model.tar.gz"), str(d / "test_project.json"), predictions_output_file=str(res), batch_size=4, cuda_device=0)
    lines = [json.loads(l) for l in res.read_text().splitlines()]
    assert [len(l) for l in lines] == [4, 4, 2]                                   # batch=4 (configs[0])
    flat = [r for l in lines for r in l]
    assert [r["Issue_Url"] for r in flat] == ["u0", "u4", "u8", "u1", "u2", "u3", "u5", "u6", "u7", "u9"]   # label groups in first-seen order
    assert all(set(r) == {"Issue_Url", "label", "predict", "prob"} for r in flat)
    tk = build_tokenizer(tok)
    by_url = {r["Issue_Url"]: r for r in rows}
    for r in flat:
        s = by_url[r["Issue_Url"]]
        ids = torch.as_tensor(tk.ids(tk.tokenize(f"{s['Issue_Title']}. {s['Issue_Body']}")), dtype=torch.int64)[None]
        ref = O.single_forward(sd, ids, torch.ones_like(ids, dtype=torch.bool), None, BERT_TINY_H512)
        assert abs(r["prob"] - float(ref["probs"][0, 1])) < TOL and r["label"] == ("pos" if s["Security_Issue_Full"] else "neg")
    assert "accuracy" in metrics and "pos_f1-score" in metrics
    m = PS.cal_metrics(str(res))
    assert m["TP"] + m["FN"] == 3 and m["TN"] + m["FP"] == 7
