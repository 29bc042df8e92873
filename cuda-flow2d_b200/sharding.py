"""Work partitioning for one-process-per-GPU runs.  No data-path collective:
frame pairs of a sequence are independent units (pair i -> rank i mod N); a single huge frame is
split into row slabs whose sizes differ by at most one row.  torch.distributed is used only to
close the timing window (barrier + max over ranks)."""


def pairs_for_rank(n_pairs, rank, world):
    """Indices of the frame pairs rank `rank` of `world` processes computes (round robin)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_pairs, world))


def slab_rows(height, rank, world):
    """[y0, y1) of the row slab of rank `rank`: contiguous, disjoint, covering, sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(height, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


def reduce_step_time(local_ms, dist=None, device=None):
    """Max over ranks of a device-timed duration (the only collective of the batch path)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms)
    import torch
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
