"""BASELINE configs[4]: mixed seq_len {128,256,512} stream at 1:300 CIR:NCIR, batch 512, sharded over the visible GPUs.
    python tools/run_config5.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_config5.py
Ranks take token-balanced shares (dist.balanced_assignment), each rank runs its share in length-bucketed batches
(collate.plan_length_buckets) of at most 64, results are gathered with one all-gather.  Prints issues/s for the naive
order (pad every batch to its longest member, as the reference does) and for bucketing."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memvul_b200.collate import plan_length_buckets
from memvul_b200.dist import balanced_assignment, gather_match
from memvul_b200.synthetic import BERT_BASE, build_memory_model, synthetic_ids

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model, _ = build_memory_model(BERT_BASE, device=dev)
g = torch.Generator().manual_seed(5)
model._golden_instances_embeddings = torch.relu(torch.randn(129, 512, generator=g)).to(dev)
model._golden_instances_labels = [f"CWE-{i}" for i in range(129)]
N = 512
lens = [[128, 256, 512][int(i)] for i in torch.randint(0, 3, (N,), generator=g)]
mine = balanced_assignment(lens, world)[rank]
counts = [len(b) for b in balanced_assignment(lens, world)]

def run(batches):
    outs = {}
    for idx in batches:
        ids_l = [mine[i] for i in idx]
        S = max(lens[i] for i in ids_l)
        ids, mask, tids = synthetic_ids(len(ids_l), S, lens=[lens[i] for i in ids_l], seed=1000 + ids_l[0])
        r = model.match_batch({"tokens": {"token_ids": ids.to(dev, non_blocking=True), "mask": mask.to(dev, non_blocking=True),
                                          "type_ids": tids.to(dev, non_blocking=True)}})
        for k, i in enumerate(idx):
            outs[i] = (r["best_idx"][k:k + 1], r["best_probs"][k:k + 1], r["probs"][k:k + 1])
    res = {"best_idx": torch.cat([outs[i][0] for i in range(len(mine))]), "best_probs": torch.cat([outs[i][1] for i in range(len(mine))]),
           "probs": torch.cat([outs[i][2] for i in range(len(mine))])}
    return gather_match(res, counts, full=True) if world > 1 else res

local_lens = [lens[i] for i in mine]
naive = [list(range(i, min(len(mine), i + 64))) for i in range(0, len(mine), 64)]
bucketed = plan_length_buckets(local_lens, 64, window=64)
with torch.no_grad():
    for name, plan in (("naive", naive), ("bucketed", bucketed)):
        run(plan)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            out = run(plan)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        dt = (time.perf_counter() - t0) / 3
        if rank == 0:
            print(f"config5 [{name}] world={world}: {N} mixed-length issue reports in {dt*1e3:.1f} ms -> {N/dt:.0f} issues/s "
                  f"(gathered probs {tuple(out['probs'].shape)})", flush=True)
if world > 1:
    dist.destroy_process_group()
