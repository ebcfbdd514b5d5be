"""Multi-GPU driver pieces for the ZEB-style sweep: image pairs are independent, so they are sharded across
ranks (one process per GPU) with NO collective on the inner loop; after the loop the ranks exchange one int64
match count each (NCCL all_gather) and, optionally, the packed result rows (variable-length gather to rank 0).

Replaces, for this path, Lightning's DistributedSampler split and the pickled-object gather over gloo
(reference: test.py:193-198, trainer/lightning.py:248-255, tools/comm.py:141-176).  Unlike the reference's sampler
the partition has no padding duplicates, so no post-hoc de-duplication by identifier is needed."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced partition of range(n_items): the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_counts(count, device=None):
    """all_gather of one int64 per rank -> list[int] (length world)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [int(count)]
    world = dist.get_world_size()
    t = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]


def gather_rows(rows, dst=0):
    """Variable-length gather of [M_r, W] row blocks to `dst` (rank order preserved).  Only `dst` allocates the
    world x max(M_r) receive buffers (dist.gather); the other ranks send their padded block and return None."""
    if not (dist.is_available() and dist.is_initialized()):
        return rows
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = gather_counts(rows.shape[0], device=rows.device)
    width = rows.shape[1]
    cap = max(counts) if counts else 0
    pad = torch.zeros(cap, width, dtype=rows.dtype, device=rows.device)
    pad[: rows.shape[0]] = rows
    bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def pack_matches(pair_ids, data):
    """[M, 6] float64 rows (pair_id, x0, y0, x1, y1, conf) from one forward's outputs; pair_ids maps batch index ->
    global pair id.  float64 keeps ids exact up to 2^53 (a float32 column would alias ids above 2^24)."""
    b = data["m_bids"].long()
    pid = torch.as_tensor(pair_ids, device=b.device, dtype=torch.float64)[b]
    return torch.cat([pid[:, None], data["mkpts0_f"].double(), data["mkpts1_f"].double(),
                      data["mconf"].double()[:, None]], 1)
