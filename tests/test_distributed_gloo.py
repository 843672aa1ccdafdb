"""The N > 1 path on CPU: world_size-2 gloo.  Ranks shard the sample range (no data-path collective), each
generates its shard -- here with the oracle standing in for the GPU kernels --, and the gathered buffers and
summed counters must equal the single-process result: results depend on (seed, global sample index) only."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from zoic_b200.distributed import shard_range  # noqa: E402


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1000, 2_123_366_400, 2_123_366_401):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, n, result_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from oracle import port
    from zoic_b200.distributed import gather_rays, reduce_stats, shard_range
    from zoic_b200.workloads import config4
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = config4()
    first, count = shard_range(n, rank, world)
    base = 3_000_000_000  # a window in the middle of the 7680x4320x128 grid
    s = port.synth_samples(wl.W, wl.H, wl.spp, wl.seed, base + first, count)
    cam = port.PortCamera(**wl.params)
    o, d, st = cam.generate(s, seed=wl.seed, first_index=base + first)
    g = gather_rays(torch.from_numpy(np.concatenate([o, d], axis=1)))
    total = reduce_stats(st, torch.device("cpu"))
    if rank == 0:
        np.savez(result_path, o=g.numpy()[:, :4], d=g.numpy()[:, 4:], stats=np.array([total[k] for k in sorted(total)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4001, 4000], ids=["ragged", "equal"])   # odd: the shards differ by one sample; even: the one-collective path
def test_two_rank_sharding_and_gather_match_single_process(tmp_path, n):
    import torch.multiprocessing as mp
    from oracle import port
    from zoic_b200.workloads import config4
    from zutil import bits_equal
    world = 2
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    got = np.load(out)
    wl = config4()
    base = 3_000_000_000
    s = port.synth_samples(wl.W, wl.H, wl.spp, wl.seed, base, n)
    cam = port.PortCamera(**wl.params)
    o, d, st = cam.generate(s, seed=wl.seed, first_index=base)
    assert bits_equal(got["o"], o) and bits_equal(got["d"], d)
    assert list(got["stats"]) == [st[k] for k in sorted(st)]


def test_pass_major_shards_render_the_whole_film():
    """bench.py's layout: rank r owns samples [r*n, (r+1)*n) of a W x H x spp job repeated
    `world` times, n = W*H*spp.  The pixel of sample i wraps per pass, so every rank covers every pixel with its
    own spp samples -- the same work mix on every rank -- and no two ranks share a sample."""
    from oracle import port
    port.load()
    W, H, spp, world, seed = 12, 8, 4, 3, 77
    n = W * H * spp
    shards = [port.synth_samples(W, H, spp, seed, r * n, n) for r in range(world)]
    for s in shards:
        px = np.floor((s[:, 0] + 1.0) * 0.5 * W).astype(int)
        py = np.floor((1.0 - s[:, 1] * (W / H)) * 0.5 * H).astype(int)
        counts = np.zeros((H, W), int)
        np.add.at(counts, (np.clip(py, 0, H - 1), np.clip(px, 0, W - 1)), 1)
        assert (counts == spp).all()          # every pixel, spp samples each
    for a in range(world):
        for b in range(a + 1, world):
            assert not np.array_equal(shards[a], shards[b])
            assert len(set(map(bytes, shards[a])) & set(map(bytes, shards[b]))) == 0


def _tile_worker(rank, world, port_no, result_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from zoic_b200.distributed import TileGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, width, tiles = 257, 8, 5
    g = TileGather(rows, width, torch.float32, torch.device("cpu"))
    bufs = [torch.empty((rows, width)), torch.empty((rows, width))]
    seen = []
    for k in range(tiles):
        b = k & 1
        if k >= 2:
            seen.append(g.wait(b).clone())   # tile k-2, gathered while tile k-1 was being produced
        bufs[b].copy_(torch.arange(rows * width, dtype=torch.float32).reshape(rows, width) + 1000.0 * k + 100000.0 * rank)
        g.submit(b, bufs[b])
    order = [(tiles - 2) & 1, (tiles - 1) & 1]
    for b in order:
        seen.append(g.wait(b).clone())
    if rank == 0:
        np.save(result_path, torch.stack(seen).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_double_buffered_tile_gather(tmp_path):
    """TileGather: every tile arrives complete and in rank order although the next tile is written while it travels."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "tiles.npy")
    world = 2
    mp.spawn(_tile_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    rows, width, tiles = 257, 8, 5
    assert got.shape == (tiles, world * rows, width)
    base = np.arange(rows * width, dtype=np.float32).reshape(rows, width)
    for k in range(tiles):
        for r in range(world):
            assert np.array_equal(got[k, r * rows:(r + 1) * rows], base + 1000.0 * k + 100000.0 * r)
