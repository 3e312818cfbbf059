"""Batch sharding across ranks: instances are independent (no reduction across instances anywhere in
the reference's solve, src/eicos.cpp:848-1262), so rank r owns one contiguous slice and the only
exchange is an optional gather of results / a max-reduction of timings."""


def shard_range(batch, rank, world):
    """[lo, hi) of rank `rank`: contiguous, sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(int(batch), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_exit_flags(local_flags, batch, group=None):
    """All ranks receive the full exit-flag vector (torch.distributed must be initialised)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    pad = max(sizes) if sizes else 0
    t = torch.full((pad,), -9999, dtype=torch.int32)
    t[:sizes[rank]] = torch.as_tensor(local_flags, dtype=torch.int32)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)]).numpy()


def max_over_ranks(value, device=None, group=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
