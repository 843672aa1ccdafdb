"""Sharding of a sample range over ranks and the final gather of the ray buffer.

Every sample is independent (per-sample retry streams), so ranks take contiguous index ranges and there is no
data-path collective; the only collective is the optional gather of the finished ray buffers and a sum of the
counters.  torch.distributed is the plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""


def shard_range(n, rank, world):
    """Contiguous, balanced split of sample indices [0, n): returns (first, count) of `rank`.
    The first n % world ranks get one extra sample; concatenating the shards in rank order gives [0, n)."""
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_rays(rays, group=None):
    """All-gather the per-rank ray buffers ([count, 8], one 32-byte record per row) into the full [n, 8] buffer,
    in rank order.  Shards may differ by one row (shard_range), so rows are padded to the largest shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=rays.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([rays.shape[0]], dtype=torch.int64, device=rays.device), group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    width = rays.shape[1]
    if min(counts) == m:
        # equal shards (the usual case): one collective straight into the final buffer -- no per-rank staging tensors and
        # no concatenation pass over the gathered records
        full = torch.empty((world * m, width), dtype=rays.dtype, device=rays.device)
        dist.all_gather_into_tensor(full, rays.contiguous(), group=group)
        return full
    pad = rays if rays.shape[0] == m else torch.cat([rays, rays.new_zeros((m - rays.shape[0], width))])
    out = [torch.empty((m, width), dtype=rays.dtype, device=rays.device) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([out[r][:counts[r]] for r in range(world)])


def reduce_stats(stats, device, group=None):
    """Sum the per-rank counter dicts (zoicb_stats) over all ranks."""
    import torch
    import torch.distributed as dist
    keys = sorted(stats)
    t = torch.tensor([stats[k] for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


class TileGather:
    """Double-buffered all-gather of finished ray tiles: the gather of tile k runs on the collective's
    own stream / thread while tile k+1 is generated into the other tile buffer.  All ranks submit tiles of the same
    shape.  Usage, with two tile buffers t[0], t[1]:

        g = TileGather(rows, width, dtype, device)
        for k in range(tiles):
            b = k & 1
            g.wait(b)                    # the gather that last read t[b] has finished: t[b] may be overwritten
            generate(out=t[b])
            g.submit(b, t[b])            # asynchronous; g.wait(b) later returns the gathered [world * rows, width] tile
        g.drain()
    """

    def __init__(self, rows, width, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.full = [torch.empty((self.world * rows, width), dtype=dtype, device=device) for _ in range(2)]
        self.work = [None, None]

    def submit(self, b, tile):
        import torch.distributed as dist
        assert self.work[b] is None, "wait(b) before re-using buffer b"
        self.work[b] = dist.all_gather_into_tensor(self.full[b], tile.contiguous(), group=self.group, async_op=True)

    def wait(self, b):
        """Blocks (stream-orders, on CUDA) until the gather submitted for buffer b is complete; returns the gathered tile."""
        if self.work[b] is not None:
            self.work[b].wait()
            self.work[b] = None
        return self.full[b]

    def drain(self):
        return [self.wait(0), self.wait(1)]
