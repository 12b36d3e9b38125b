"""Multi-process (world_size 2, gloo, CPU) tests of the batch-sharding host logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from memvul_b200.dist import balanced_assignment, build_memory_sharded, gather_match, gather_rows, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 1024, 1025):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_assignment_mixed_lengths():
    g = torch.Generator().manual_seed(0)
    lens = [int([128, 256, 512][i]) for i in torch.randint(0, 3, (512,), generator=g)]
    buckets = balanced_assignment(lens, 8)
    assert sorted(i for b in buckets for i in b) == list(range(512))
    cost = lambda s: 14155776.0 * s + 3072.0 * s * s
    loads = [sum(cost(lens[i]) for i in b) for b in buckets]
    assert max(loads) / min(loads) < 1.02


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G = 5
        counts = [3, 2]                                  # ragged shards: B = 5 over 2 ranks
        start = sum(counts[:rank])
        b = counts[rank]
        full_probs = torch.arange(5 * G * 2, dtype=torch.float32).view(5, G, 2) / 100.0
        local = {"probs": full_probs[start:start + b].clone(),
                 "best_idx": torch.tensor([(start + i) % G for i in range(b)], dtype=torch.int32),
                 "best_probs": full_probs[start:start + b, 0].clone()}
        out = gather_match(local, counts, full=True)
        assert torch.equal(out["probs"], full_probs)
        assert out["best_idx"].tolist() == [i % G for i in range(5)]
        assert torch.equal(out["best_probs"], full_probs[:, 0])
        red = gather_match(local, counts, full=False)
        assert "probs" not in red and red["best_idx"].tolist() == [i % G for i in range(5)]
        rows = gather_rows(torch.full((counts[rank], 4), float(rank)), counts)
        assert rows.shape == (5, 4) and rows[:, 0].tolist() == [0.0, 0.0, 0.0, 1.0, 1.0]
        eq = gather_rows(torch.full((2, 3), float(rank)), [2, 2])
        assert eq[:, 0].tolist() == [0.0, 0.0, 1.0, 1.0]
        with pytest.raises(ValueError):
            gather_rows(torch.zeros(1, 2), [3, 3])
    finally:
        dist.destroy_process_group()


def test_gather_match_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)


class _StubMemoryModel:
    """The two attributes and one method of ModelMemory that the sharded bank build touches; 'encoding' an anchor is a
    deterministic function of the anchor alone, like the real encoder."""

    def __init__(self):
        self._golden_instances_embeddings = None
        self._golden_instances_labels = None
        self.calls = []

    def forward_on_instances(self, instances):
        self.calls.append(len(instances))
        rows = torch.stack([torch.full((4,), float(i["id"])) + torch.arange(4.0) / 10 for i in instances])
        labels = [i["label"] for i in instances]
        if self._golden_instances_embeddings is None:
            self._golden_instances_embeddings, self._golden_instances_labels = rows, labels
        else:
            self._golden_instances_embeddings = torch.cat([self._golden_instances_embeddings, rows])
            self._golden_instances_labels = self._golden_instances_labels + labels


def _bank_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        golden = [{"id": i, "label": f"CWE-{i % 5}"} for i in range(11)]            # ragged: 6 + 5 anchors
        single = _StubMemoryModel()
        single.forward_on_instances(golden)
        m = _StubMemoryModel()
        m._golden_instances_embeddings = torch.zeros(3, 4)                              # a stale bank must be replaced
        build_memory_sharded(m, golden, chunk=4)
        assert torch.equal(m._golden_instances_embeddings, single._golden_instances_embeddings)
        assert m._golden_instances_labels == single._golden_instances_labels
        assert m.calls == ([4, 2] if rank == 0 else [4, 1])                             # own slice only, in chunks
        with pytest.raises(ValueError):
            build_memory_sharded(_StubMemoryModel(), golden[:1])
    finally:
        dist.destroy_process_group()


def test_build_memory_sharded_world2_gloo():
    """rank-sharded anchor encoding + one all-gather of the bank rows == the single-process bank."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_bank_worker, args=(2, port), nprocs=2, join=True)
