"""Multi-GPU plumbing: replicates / prior draws shard across ranks with no data-path collective;
the only exchange is the all-gather of the accepted ABC draws at the end.

One process per GPU.  The reference spreads idx in [seed*10, seed*10+runs) over rayon workers
(src/main.rs:214-225); here the same index range is cut into contiguous blocks, one per rank."""


def rank_range(idx_begin, n_runs, rank, world):
    """Contiguous block of the replicate index range owned by `rank` (first ranks get the remainder)."""
    base, rem = divmod(n_runs, world)
    start = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return idx_begin + start, count


def gather_accepted(torch, dist, payload):
    """All-gather a [n_accepted, C] tensor whose first dimension differs per rank, through torch.distributed: for
    jobs whose ranks hold accepted records as tensors (e.g. a CPU post-processing step over gloo).  The GPU path of
    the library is ecdna_b200_abc_pack + ecdna_b200_abc_allgather (csrc/abc_gather.cu), which needs no host round trip.

    Two collectives: the counts, then fixed-stride records padded to the largest count (NCCL has no
    all-gather-v).  Returns the concatenation in rank order."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return payload
    n = torch.tensor([payload.shape[0]], dtype=torch.int64, device=payload.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(payload.shape[1:]), dtype=payload.dtype, device=payload.device)
    pad[: payload.shape[0]] = payload
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
