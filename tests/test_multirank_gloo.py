"""world_size-2 `gloo` test of the multi-GPU host logic (CPU): index sharding and the one collective
of the path, the all-gather of accepted ABC draws with per-rank counts."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_range_partitions_the_index_range(pkg):
    for n, world in [(10, 2), (10000, 8), (7, 4), (3, 8), (1000001, 8)]:
        seen = []
        for r in range(world):
            b, c = pkg.rank_range(260, n, r, world)
            seen += list(range(b, b + c)) if n < 100000 else [(b, c)]
        if n < 100000:
            assert seen == list(range(260, 260 + n))
        else:
            assert sum(c for _, c in seen) == n and seen[0][0] == 260
            assert all(seen[i][0] + seen[i][1] == seen[i + 1][0] for i in range(world - 1))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import _pkg
    m = _pkg.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank "accepts" a different number of draws from its own shard of the index range
    begin, count = m.rank_range(260, 101, rank, world)
    idx = torch.arange(begin, begin + count, dtype=torch.float32)
    accepted = idx[(idx.long() % (3 + rank)) == 0]
    payload = torch.stack([accepted, accepted * 2 + rank], dim=1)  # [n_acc, 2]
    got = m.gather_accepted(torch, dist, payload)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), got.numpy())
    # the bench's timing reduction: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    dist.destroy_process_group()


def test_gather_accepted_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(a, b)  # every rank holds the full accepted set
    r0 = [i for i in range(260, 311) if i % 3 == 0]
    r1 = [i for i in range(311, 361) if i % 4 == 0]
    want = np.array([[i, 2 * i] for i in r0] + [[i, 2 * i + 1] for i in r1], dtype=np.float32)
    np.testing.assert_array_equal(a, want)


def test_merge_and_decode_gathered_records(pkg):
    """What ecdna_b200_abc_allgather leaves on every rank ([world][capacity] record blocks + [world] counts) merged
    in rank order and decoded by the layout of include/ecdna_b200.h; an overflowing rank is an error, not a
    silent truncation."""
    import pytest
    bins, cap, world = 8, 5, 3
    words = pkg.record_words(bins)
    assert words == pkg.REC_HEADER + bins == 24
    blocks = np.zeros((world, cap, words), dtype=np.uint32)
    counts = np.array([2, 0, 3])
    idx = 260
    for r in range(world):
        for j in range(counts[r]):
            rec = blocks[r, j]
            rec[0], rec[1] = idx & 0xFFFFFFFF, 1  # index with a high word
            rec[2:6] = np.array([1.0, 1.5, 0.1, 0.2], dtype=np.float32).view(np.uint32)
            rec[6:10] = np.array([0.01 * (idx - 259), 0.02, 0.03, 0.04], dtype=np.float32).view(np.uint32)
            rec[10:13] = np.array([3.5, 0.9, 2.25], dtype=np.float32).view(np.uint32)
            rec[13], rec[14], rec[15] = 1000, 7, pkg.STOP_MAX_CELLS | pkg.FLAG_SPILLED
            rec[pkg.REC_HEADER:] = np.arange(bins) + idx
            idx += 1
    merged = pkg.merge_gathered(blocks, counts, cap, bins)
    assert merged.shape == (5, words)
    d = pkg.decode_records(merged, bins)
    assert list(d["idx"]) == [(1 << 32) | (260 + i) for i in range(5)]
    np.testing.assert_array_equal(d["rates"][0], np.array([1.0, 1.5, 0.1, 0.2], dtype=np.float32))
    np.testing.assert_allclose(d["distance"][:, 0], 0.01 * np.arange(1, 6), rtol=1e-6)
    assert d["mean"][0] == np.float32(3.5) and d["entropy"][4] == np.float32(2.25) and list(d["cells"]) == [1000] * 5
    assert np.all(d["stop"] == pkg.STOP_MAX_CELLS) and d["hist"][3][0] == 263 and d["hist"].shape == (5, bins)
    with pytest.raises(OverflowError):
        pkg.merge_gathered(blocks, np.array([2, 9, 3]), cap, bins)
