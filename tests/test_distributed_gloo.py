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
    s = port.synth_samples(*wl.synth_args(), base + first, count)
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
    s = port.synth_samples(*wl.synth_args(), base, n)
    cam = port.PortCamera(**wl.params)
    o, d, st = cam.generate(s, seed=wl.seed, first_index=base)
    assert bits_equal(got["o"], o) and bits_equal(got["d"], d)
    assert list(got["stats"]) == [st[k] for k in sorted(st)]


def test_pass_major_shards_render_the_whole_film():
    """The frame layout (zoic_b200.workloads): a W x H x spp frame in 8 passes of spp / 8 samples per
    pixel; rank r of G owns passes [r 8/G, (r+1) 8/G) = samples [r N/G, (r+1) N/G).  The pixel of sample i wraps per pass,
    so every rank covers every pixel with its own samples -- the same work mix on every rank --, no two ranks share a
    sample, and the union over the ranks is the same sample set for every G."""
    from oracle import port
    from zoic_b200.distributed import job_share
    port.load()
    W, H, spp, passes, seed = 12, 8, 16, 8, 77
    n = W * H * spp
    whole = port.synth_samples(W, H, spp // passes, seed, 0, n)
    for world in (1, 2, 4, 8):
        shards = []
        for r in range(world):
            first, count = job_share(n, passes, r, world)
            assert first == r * n // world and count == n // world
            shards.append(port.synth_samples(W, H, spp // passes, seed, first, count))
        for s in shards:
            px = np.floor((s[:, 0] + 1.0) * 0.5 * W).astype(int)
            py = np.floor((1.0 - s[:, 1] * (W / H)) * 0.5 * H).astype(int)
            counts = np.zeros((H, W), int)
            np.add.at(counts, (np.clip(py, 0, H - 1), np.clip(px, 0, W - 1)), 1)
            assert (counts == spp // world).all()          # every pixel, spp / world samples each
        assert np.array_equal(np.concatenate(shards), whole)   # the same job for every world size
    with pytest.raises(ValueError):
        job_share(n, passes, 0, 3)


def _share_worker(rank, world, port_no, result_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from oracle import port
    from zoic_b200.distributed import gather_rays, job_share, reduce_stats
    from zoic_b200.workloads import Workload, _kolb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = Workload("tiny double gauss", 16, 9, 16, 5, _kolb("double_gauss_f2.0.dat", 5.0, 2.0))
    first, count = job_share(wl.n, wl.passes, rank, world)
    s = port.synth_samples(*wl.synth_args(), first, count)
    cam = port.PortCamera(**wl.params)
    o, d, st = cam.generate(s, seed=wl.seed, first_index=first)
    g = gather_rays(torch.from_numpy(np.concatenate([o, d], axis=1)))
    total = reduce_stats(st, torch.device("cpu"))
    if rank == 0:
        np.savez(result_path, rays=g.numpy(), stats=np.array([total[k] for k in sorted(total)]))
    dist.barrier()
    dist.destroy_process_group()


def test_strong_split_of_a_frame_equals_the_single_process_job(tmp_path):
    """bench.py's multi-GPU job on CPU: two ranks take their job_share of one small frame (the oracle standing in for the
    kernels), the gathered records and the summed counters equal the whole frame generated by one process, bit for bit."""
    import torch.multiprocessing as mp
    from oracle import port
    from zoic_b200.workloads import Workload, _kolb
    from zutil import bits_equal
    world = 2
    out = str(tmp_path / "frame.npz")
    mp.spawn(_share_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    wl = Workload("tiny double gauss", 16, 9, 16, 5, _kolb("double_gauss_f2.0.dat", 5.0, 2.0))
    s = port.synth_samples(*wl.synth_args(), 0, wl.n)
    cam = port.PortCamera(**wl.params)
    o, d, st = cam.generate(s, seed=wl.seed, first_index=0)
    assert bits_equal(got["rays"], np.concatenate([o, d], axis=1))
    assert list(got["stats"]) == [st[k] for k in sorted(st)]
