"""Parity on the BASELINE.json configurations at their stated sizes, with non-vacuous gates (SURVEY.md 8d "parity gates
per run"; VERDICT r01 "next round" item 1):

  C2  all 64 S=512 rows against the CPU oracle (oracle-built bank), labels at 0.5 and at a splitting threshold compared
      on EVERY row, arg-max compared on the rows the probe (tools/margin_probe.py) pinned as clear;
  C4  the tiled match at 256 x 16,384 against the literal concat/Linear statement evaluated in anchor chunks;
  C5  a bert-base batch mixing {128, 256, 512} in data order (packed), padded, and length-bucketed, against the oracle;
  NCCL  the sharded path's single all-gather on 2 GPUs (skips on a 1-GPU box).
"""
import json
import os
import socket

import numpy as np
import pytest
import torch

from config_inputs import c2_inputs, c5_inputs, split_threshold

pytestmark = pytest.mark.gpu
TOL = 1e-3          # north_star: "match logits within 1e-3 absolute"


def _dev(ids, mask, tids=None):
    d = {"token_ids": ids.cuda(), "mask": mask.cuda()}
    d["type_ids"] = (torch.zeros_like(ids) if tids is None else tids).cuda()
    return {"tokens": d}


def _build_bank(model, a_ids, a_mask, chunk=128):
    for c0 in range(0, a_ids.shape[0], chunk):
        ids, mask = a_ids[c0:c0 + chunk], a_mask[c0:c0 + chunk]
        S = int(mask.sum(1).max())
        model.forward_gold_instances(_dev(ids[:, :S].contiguous(), mask[:, :S].contiguous()),
                                     [{"type": "golden", "instance": [{"label": f"CWE-{c0 + i}"}]} for i in range(ids.shape[0])])


def test_c2_all_64_rows_against_the_oracle():
    from memvul_b200.parity import gate_report
    from memvul_b200.synthetic import BERT_BASE, build_memory_model
    from oracle import memvul_oracle as O
    a_ids, a_mask, alens, ids, mask, tids, lens = c2_inputs()
    model, sd = build_memory_model(BERT_BASE, device="cuda")
    with torch.no_grad():
        _build_bank(model, a_ids, a_mask)                                   # 128 + 1 chunks (predict_memory.py:81-83)
        res = model.match_batch(_dev(ids, mask, tids))
        bank = O.build_bank(sd, [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(len(alens))])   # ORACLE-built bank
        ref = O.memory_forward(sd, ids, mask, tids, bank, model._same_idx)
    bank_err = float((model._golden_instances_embeddings.cpu() - bank).abs().max())
    vote_ref = ref["p"][:, :, model._same_idx].max(1).values
    thr, margin = split_threshold(vote_ref)
    rep = gate_report(res["logits"].cpu().numpy(), res["probs"].cpu().numpy(), res["best_idx"].cpu().numpy(),
                      ref["logits"].numpy(), ref["p"].numpy(), model._same_idx, thresholds=(0.5, thr), tol=TOL)
    rep["bank_err"], rep["split_threshold"], rep["split_margin"] = bank_err, thr, margin
    print("C2 parity gates:", json.dumps(rep))
    assert bank_err < TOL
    assert rep["max_logit_err"] <= TOL, rep
    # the probe pinned these properties of the ORACLE outputs, so none of the gates below can pass vacuously
    assert rep["labels"]["0.5"]["min_margin"] > 5e-3 and 0 < rep["labels"][f"{thr:g}"]["pos_ref"] < 64
    assert rep["argmax_clear_rows"] >= 60
    # labels: identical on EVERY row at both thresholds (not only outside the tolerance band)
    assert all(v["mismatch_rows"] == 0 for v in rep["labels"].values()), rep
    assert rep["argmax_mismatch_clear"] == 0 and rep["argmax_not_maximiser"] == 0, rep
    assert rep["ok"]


@pytest.mark.parametrize("B,G,same", [(256, 16384, 0), (100, 4100, 1)])
def test_c4_tiled_match_against_the_literal_statement(B, G, same):
    """BASELINE configs[3]: 256 queries x 16,384 anchors goes through the shared-memory tiled match; compared with the
    reference's expand / cat / abs / Linear / softmax / argmax (model_memory.py:133-147) evaluated in anchor chunks
    (the full [B,G,1536] concat is 25.8 GB); (100, 4100) exercises ragged query and anchor tiles."""
    from memvul_b200 import native as N
    from oracle import memvul_oracle as O
    H, D = 768, 512
    g = torch.Generator().manual_seed(B + G)
    cls = torch.randn(B, H, generator=g)
    wp, bp = torch.randn(H, H, generator=g) * 0.03, torch.randn(H, generator=g) * 0.02
    wh, bh = torch.randn(D, H, generator=g) * 0.03, torch.randn(D, generator=g) * 0.02
    wproj = torch.randn(2, 3 * D, generator=g) * 0.03
    bank = torch.relu(torch.randn(G, D, generator=g) * 0.4)
    c = lambda t: t.cuda().contiguous()
    bankd, wprojd = c(bank), c(wproj)
    out = N.pool_match(c(cls), H, B, c(wp), c(bp), c(wh), c(bh), wprojd, bankd, N.bank_prepare(bankd, wprojd), same_idx=same)
    lin = torch.nn.functional.linear
    u_ref = torch.relu(lin(torch.tanh(lin(cls, wp, bp)), wh, bh))
    logits, probs = out["logits"].cpu(), out["probs"].cpu()
    ref_ps = torch.empty(B, G)
    err_l = err_p = 0.0
    for g0 in range(0, G, 512):                                      # O.match materialises [B, chunk, 1536]
        r = O.match(u_ref, bank[g0:g0 + 512], wproj, same)
        err_l = max(err_l, float((logits[:, g0:g0 + 512] - r["logits"]).abs().max()))
        err_p = max(err_p, float((probs[:, g0:g0 + 512] - r["p"]).abs().max()))
        ref_ps[:, g0:g0 + 512] = r["p"][:, :, same]
    assert float((out["u"].cpu() - u_ref).abs().max()) < 2e-5
    assert err_l < 5e-5 and err_p < 2e-5, (err_l, err_p)
    idx = out["best_idx"].cpu().long()
    # the kernel's arg-max is the first maximum of its own probabilities, bit for bit ...
    assert torch.equal(idx, probs[:, :, same].argmax(1))
    assert torch.equal(out["best_probs"].cpu(), probs[torch.arange(B), idx])
    # ... and a maximiser of the reference's up to the fp32 noise between the two evaluation orders
    assert float((ref_ps.max(1).values - ref_ps[torch.arange(B), idx]).max()) < 4e-5
    top2 = ref_ps.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert int(clear.sum()) >= B // 2 and torch.equal(idx[clear], ref_ps.argmax(1)[clear])


def test_c5_mixed_lengths_packed_padded_and_bucketed():
    """BASELINE configs[4] shape at bert-base: one batch mixing {128, 256, 512} (plus ragged neighbours) through
    (i) data order with the packed var-len execution, (ii) data order padded to 512, (iii) length buckets
    (collate.plan_length_buckets) -- all three against the CPU oracle."""
    from memvul_b200 import custom_PTM_embedder as E
    from memvul_b200.collate import plan_length_buckets
    from memvul_b200.parity import gate_report
    from memvul_b200.synthetic import BERT_BASE, build_memory_model
    from oracle import memvul_oracle as O
    a_ids, a_mask, alens, ids, mask, tids, lens = c5_inputs()
    model, sd = build_memory_model(BERT_BASE, device="cuda")
    same = model._same_idx
    with torch.no_grad():
        _build_bank(model, a_ids, a_mask)
        bank = O.build_bank(sd, [(a_ids[i][a_mask[i]], a_mask[i][a_mask[i]]) for i in range(len(alens))])
        ref = O.memory_forward(sd, ids, mask, tids, bank, same)
        assert E._PACKED_DEFAULT
        packed = {k: v.clone() for k, v in model.match_batch(_dev(ids, mask, tids)).items() if k in ("logits", "probs", "best_idx")}
        E._PACKED_DEFAULT = False
        try:
            padded = {k: v.clone() for k, v in model.match_batch(_dev(ids, mask, tids)).items() if k in ("logits", "probs", "best_idx")}
        finally:
            E._PACKED_DEFAULT = True
        buck = {k: torch.empty_like(v) for k, v in packed.items()}
        for idx in plan_length_buckets(lens, batch_size=4, window=3):
            S_b = max(lens[i] for i in idx)
            r = model.match_batch(_dev(ids[idx][:, :S_b].contiguous(), mask[idx][:, :S_b].contiguous()))
            for k in buck:
                buck[k][torch.tensor(idx, device="cuda")] = r[k]
    vote_ref = ref["p"][:, :, same].max(1).values
    thr, margin = split_threshold(vote_ref)
    for name, got in (("packed", packed), ("padded", padded), ("bucketed", buck)):
        rep = gate_report(got["logits"].cpu().numpy(), got["probs"].cpu().numpy(), got["best_idx"].cpu().numpy(),
                          ref["logits"].numpy(), ref["p"].numpy(), same, thresholds=(0.5, thr), tol=TOL)
        print(f"C5 {name}: max_logit_err {rep['max_logit_err']:.2e} split margin {margin:.2e} clear {rep['argmax_clear_rows']}/12 "
              f"mismatch {[v['mismatch_rows'] for v in rep['labels'].values()]}")
        assert rep["max_logit_err"] <= TOL, (name, rep)
        assert all(v["mismatch_outside_tol"] == 0 for v in rep["labels"].values()), (name, rep)
        assert rep["argmax_mismatch_clear"] == 0 and rep["argmax_not_maximiser"] == 0, (name, rep)
    assert 0 < int((vote_ref >= thr).sum()) < len(lens)              # the split threshold really splits
    # the packed and the padded execution run the same kernels on the same rows: not a single bit may change
    assert torch.equal(packed["logits"], padded["logits"]), float((packed["logits"] - padded["logits"]).abs().max())
    # (small buckets take the single-CTA GEMM / stand-alone LayerNorm kernels: same math, different rounding order)
    assert float((packed["logits"] - buck["logits"]).abs().max()) < TOL / 2


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from memvul_b200.dist import AsyncGather, gather_match, shard_bounds
        from memvul_b200.synthetic import BERT_TINY, build_memory_model, synthetic_ids
        dev = torch.device("cuda", rank)
        model, _ = build_memory_model(BERT_TINY, device=dev)
        alens = [9, 33, 64, 17, 40]
        a_ids, a_mask, _ = synthetic_ids(5, 64, lens=alens, seed=3, vocab_size=1024)
        lens = [64, 20, 3, 50, 64, 31, 7]                                        # 7 reports over 2 ranks: ragged shards
        ids, mask, tids = synthetic_ids(7, 64, lens=lens, seed=4, vocab_size=1024)
        to = lambda t: t.to(dev)
        with torch.no_grad():
            model.forward_gold_instances({"tokens": {"token_ids": to(a_ids), "mask": to(a_mask), "type_ids": to(torch.zeros_like(a_ids))}},
                                         [{"type": "golden", "instance": [{"label": f"CWE-{i}"}]} for i in range(5)])
            whole = model.match_batch({"tokens": {"token_ids": to(ids), "mask": to(mask), "type_ids": to(tids)}})
            bounds = shard_bounds(7, world)
            counts = [e - s for s, e in bounds]
            s, e = bounds[rank]
            sh = {"tokens": {"token_ids": to(ids[s:e].contiguous()), "mask": to(mask[s:e].contiguous()), "type_ids": to(tids[s:e].contiguous())}}
            ag = AsyncGather(counts, dev)
            for _ in range(3):                                                   # the collective overlaps the next batch
                ag.submit(model.match_batch(sh, flat_capacity=max(counts)))
            parts = ag.wait()
            torch.cuda.synchronize()
            got = {k: torch.cat(v) for k, v in parts.items()}
            old = gather_match(model.match_batch(sh), counts, full=True)          # the packed-copy form gives the same
            # rank-sharded bank build (each rank encodes its slice, ONE all-gather of the rows) == the bank every
            # rank built for itself above (same chunking inside the slices: 5 anchors -> 3 + 2)
            from memvul_b200.dist import build_memory_sharded
            bank_full, labels_full = model._golden_instances_embeddings.clone(), list(model._golden_instances_labels)
            golden = [{"sample1": {"token_ids": a_ids[i][:alens[i]].tolist(), "type_ids": [0] * alens[i]}, "label": None,
                       "metadata": {"type": "golden", "instance": [{"label": f"CWE-{i}"}]}} for i in range(5)]
            build_memory_sharded(model, golden)
            bank_ok = (model._golden_instances_labels == labels_full
                       and float((model._golden_instances_embeddings - bank_full).abs().max()) < 1e-5)
            again = model.match_batch({"tokens": {"token_ids": to(ids), "mask": to(mask), "type_ids": to(tids)}})
            bank_ok = bank_ok and float((again["logits"] - whole["logits"]).abs().max()) < 1e-5
        ok = (torch.equal(got["probs"], whole["probs"]) and torch.equal(got["best_idx"], whole["best_idx"])
              and torch.equal(got["best_probs"], whole["best_probs"]) and torch.equal(old["probs"], whole["probs"])
              and torch.equal(old["best_idx"], whole["best_idx"]) and bank_ok)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_nccl_gather_match_two_ranks():
    """SURVEY 8e on real GPUs: shard results travel in ONE all_gather_into_tensor (NCCL) on a side stream and every rank
    ends up with exactly the single-GPU result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}
