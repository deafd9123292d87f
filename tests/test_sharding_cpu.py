"""Host logic of the N > 1 path on CPU: world_size-2 gloo process groups."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cdfo_b200 import sharding, synthetic


def test_partition_covers_everything_once():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.partition(n, world, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [sharding.partition(n, world, r)[1] - sharding.partition(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.sequence_shard(64, 8, 3) == list(range(24, 32))      # config c4: blocks of 8 sequences on 8 GPUs


def test_frame_shard_halo_and_windows():
    (lo, hi), (rlo, rhi) = sharding.frame_shard(32, 4, 1)
    assert (lo, hi) == (8, 16) and (rlo, rhi) == (5, 19)
    (lo, hi), (rlo, rhi) = sharding.frame_shard(32, 4, 0)
    assert (rlo, rhi) == (0, 11)
    # every window of an owned frame lies inside the range the shard reads
    for world in (1, 2, 4):
        for r in range(world):
            (lo, hi), (rlo, rhi) = sharding.frame_shard(32, world, r)
            for i in range(lo, hi):
                w = sharding.window_indices(i, 32)
                assert len(w) == 7 and min(w) >= rlo and max(w) < rhi
    assert sharding.window_indices(0, 32) == [0, 0, 0, 0, 1, 2, 3]       # clipped like test_LD_37.py:13-16
    assert sharding.window_indices(31, 32) == [28, 29, 30, 31, 31, 31, 31]


def test_noise_is_keyed_by_work_item():
    a = synthetic.gumbel_uniforms(4, 5, 9, 1, 8, 8)
    b = synthetic.gumbel_uniforms(4, 5, 9, 1, 8, 8)
    c = synthetic.gumbel_uniforms(4, 6, 9, 1, 8, 8)
    assert all(torch.equal(x, y) for x, y in zip(a, b)) and not torch.equal(a[0], c[0])
    g = torch.Generator().manual_seed(sharding.noise_key(4, 5, 9, 2))
    assert torch.equal(torch.rand(1, 64, 8, 8, generator=g).clamp_min(1e-12), a[2])


def _fake_frame_metrics(seq, frame):
    """Stand-in for (PSNR dB, SSIM, 1) of one output frame -- the triple metrics.psnr_ssim(accum=) adds per frame: a deterministic
    function of the work item."""
    g = torch.Generator().manual_seed(1000 * seq + frame)
    err = torch.randn(16, 16, generator=g, dtype=torch.float64)
    return torch.tensor([float(30.0 + (err * err).mean()), float(err.abs().mean()), 1.0], dtype=torch.float64)


def _worker(rank, world, port, n_seq, n_frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = torch.zeros(n_seq, 3, dtype=torch.float64)
        for s in sharding.sequence_shard(n_seq, world, rank):
            for f in range(n_frames):
                local[s] += _fake_frame_metrics(s, f)
        total = sharding.gather_metrics(local)
        dist.barrier()
        if rank == 0:
            torch.save(total, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_metric_gather_world2_gloo(tmp_path):
    """Two ranks own disjoint sequence blocks; the all-reduced sums equal the single-process result exactly
    (each row has exactly one non-zero contribution, so there is no summation-order effect)."""
    n_seq, n_frames = 6, 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "total.pt")
    mp.spawn(_worker, args=(2, port, n_seq, n_frames, out), nprocs=2, join=True)
    total = torch.load(out)
    ref = torch.zeros(n_seq, 3, dtype=torch.float64)
    for s in range(n_seq):
        for f in range(n_frames):
            ref[s] += _fake_frame_metrics(s, f)
    assert torch.equal(total, ref)
    psnr, ssim = sharding.mean_metrics_per_sequence(total)
    assert psnr.shape == (n_seq,) and torch.allclose(psnr, ref[:, 0] / n_frames) and torch.allclose(ssim, ref[:, 1] / n_frames)


def test_gather_metrics_without_group_is_identity():
    t = torch.ones(4, 3, dtype=torch.float64)
    assert torch.equal(sharding.gather_metrics(t.clone()), t)
    with pytest.raises(ValueError):
        sharding.gather_metrics(torch.ones(4, 3))
