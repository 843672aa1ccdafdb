"""Sharding of a sample range over ranks and the final gather of the ray buffer.

Every sample is independent (per-sample retry streams), so ranks take contiguous index ranges and there is no
data-path collective; the only exchange is the final gather of the finished ray tiles to the consumer rank, which is
libzoicb's own (zoic_b200.Gather / csrc/gather.cu: generate kernels storing into the consumer's memory over NVLink, a
copy-engine push, or ncclSend/ncclRecv).  torch.distributed is plumbing here: it carries the set-up handshake
(connect_gather), the counters (reduce_stats) and, in the CPU tests, a gloo all-gather standing in for NVLink.
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


def job_share(n, passes, rank, world):
    """Rank `rank`'s share of an n-sample frame laid out in `passes` passes (zoic_b200.workloads): whole passes, i.e. the
    contiguous index range [rank n / world, (rank + 1) n / world).  Every rank renders every pixel of the film with its
    own samples, so all ranks carry the same mix of vignetted and clear pixels (bands of pixel rows lost 11 % to
    imbalance, profiles/r01c_bench_headline_8gpu_bands.json), and the rays of the job do not depend on `world`."""
    if passes % world or n % passes:
        raise ValueError("world size %d does not divide the frame's %d passes" % (world, passes))
    per = (n // passes) * (passes // world)
    return rank * per, per


def connect_gather(gather, group=None):
    """Collective set-up of a zoic_b200.Gather: exchanges the ranks' IPC blobs (and, for the NCCL transport, a unique id
    made on rank 0) over torch.distributed -- plumbing only; the data path is libzoicb's."""
    import torch.distributed as dist
    from .camera import nccl_unique_id
    world = dist.get_world_size(group)
    blobs = [None] * world
    dist.all_gather_object(blobs, gather.export(), group=group)
    gather.connect(blobs)
    if gather.transport == "nccl":
        ident = [nccl_unique_id() if dist.get_rank(group) == 0 else None]
        dist.broadcast_object_list(ident, src=0, group=group)
        gather.init_nccl(ident[0])
    dist.barrier(group=group)
