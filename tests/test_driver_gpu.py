"""GPU parity of the frame I/O / metric kernels (csrc/metrics.cu) against the oracle and the reference's golden numbers, and
of the frame driver against a plain per-frame loop that rebuilds every window on the host like test_LD_37.py does."""
import os

import numpy as np
import pytest
import torch

import golden_util as G
from cdfo_b200 import driver, metrics, sharding
from oracle import metrics_ref as R, priors_ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_psnr_ssim_matches_reference_golden(cuda_dev):
    g = np.load(os.path.join(GOLD, "metrics_golden.npz"))
    for i in range(int(g["n"])):
        res, gt = torch.from_numpy(g["res%d" % i]).to(cuda_dev), torch.from_numpy(g["gt%d" % i]).to(cuda_dev)
        out = metrics.psnr_ssim(res[None], gt[None]).cpu().numpy()[0]
        if np.isinf(g["psnr%d" % i]):
            assert np.isinf(out[0]) and abs(out[1] - 1.0) < 1e-12
        else:
            assert abs(out[0] - float(g["psnr%d" % i])) < 2e-5 and abs(out[1] - float(g["ssim%d" % i])) < 1e-10, (i, out)


@pytest.mark.parametrize("shape", [(3, 270 * 4, 480 * 4), (2, 97, 131), (1, 19, 19)])
def test_psnr_ssim_batched_vs_oracle(cuda_dev, shape):
    rng = np.random.default_rng(shape[1])
    B, H, W = shape
    gt = rng.integers(0, 256, shape, dtype=np.uint8)
    gt = (gt // 8 + np.linspace(0, 200, W).astype(np.int64)[None, None, :]).clip(0, 255).astype(np.uint8)
    res = np.clip(gt.astype(np.int64) + rng.integers(-9, 10, shape), 0, 255).astype(np.uint8)
    acc = torch.zeros((B, 3), dtype=torch.float64, device=cuda_dev)
    a, b = torch.from_numpy(res).to(cuda_dev), torch.from_numpy(gt).to(cuda_dev)
    out1 = metrics.psnr_ssim(a, b, accum=acc)
    out2 = metrics.psnr_ssim(a, b, accum=acc)
    assert torch.equal(out1, out2)                                  # fixed summation order: bit-reproducible
    out = out1.cpu().numpy()
    n_ref = B if H < 500 else 1                                     # the pure-numpy oracle takes seconds per HR frame
    for k in range(n_ref):
        assert abs(out[k, 0] - R.psnr_y(res[k], gt[k])) < 1e-9 and abs(out[k, 1] - R.ssim_y(res[k], gt[k])) < 1e-10
    acc = acc.cpu().numpy()
    assert np.allclose(acc[:, 0], 2 * out[:, 0]) and np.allclose(acc[:, 1], 2 * out[:, 1]) and np.all(acc[:, 2] == 2)


def test_psnr_ssim_errors(cuda_dev):
    a = torch.zeros((1, 16, 16), dtype=torch.uint8, device=cuda_dev)
    with pytest.raises(RuntimeError):
        metrics.psnr_ssim(a, a)                                     # 8x8 after the border crop: no room for the 11x11 window
    with pytest.raises(RuntimeError):
        metrics.psnr_ssim(a, a[:, :8])
    with pytest.raises(NotImplementedError):
        metrics.psnr_ssim(a.cpu(), a.cpu())


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.int16, np.int32])
def test_planes_to_unit_bit_exact(cuda_dev, dtype):
    rng = np.random.default_rng(1)
    lo, hi = (0, 256) if dtype == np.uint8 else (-128, 128)
    x = rng.integers(lo, hi, (2, 3, 270 // 9, 48)).astype(dtype)
    got = metrics.planes_to_unit(torch.from_numpy(x).to(cuda_dev), 32).cpu().numpy()
    assert got.shape == (2, 3, 32, 48) and np.array_equal(got, R.planes_to_unit(x, 32))


@pytest.mark.parametrize("W", [64, 50])
def test_sr_to_u8_bit_exact(cuda_dev, W):
    g = torch.Generator().manual_seed(W)
    sr = torch.rand((2, 1, 40, W), generator=g) * 1.4 - 0.2
    sr[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 254.999 / 255.0, float("nan")])
    got = metrics.sr_to_u8(sr.to(cuda_dev), 36).cpu().numpy()
    ref = R.sr_to_u8(np.nan_to_num(sr.numpy(), nan=0.0), 36)
    assert got.shape == (2, 1, 36, W) and np.array_equal(got, ref)


# ------------------------------------------------------------------------------------------------ the driver
def _sequence(seed, T, h, W, with_gt=True):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:W]
    lr = np.stack([np.clip(120 + 80 * np.sin((xx + 2 * t) / 5.0) * np.cos(yy / 4.0) + rng.normal(0, 6, (h, W)), 0, 255) for t in range(T)])
    lr = lr.astype(np.uint8)
    pm = (rng.integers(0, 4, (T, h // 8 + 1, W // 8 + 1)) * 85).astype(np.uint8).repeat(8, 1).repeat(8, 2)[:, :h, :W]
    res = np.clip(np.round(rng.normal(0, 6, (T, h, W))), -128, 127).astype(np.int16)
    unflt = np.clip(lr.astype(np.int64) + rng.integers(-4, 5, lr.shape), 0, 255).astype(np.uint8)
    mv = np.zeros((T, h, W, 3), np.int8)
    blk = rng.integers(-64, 64, (T, h // 8 + 1, W // 8 + 1, 2)).repeat(8, 1).repeat(8, 2)[:, :h, :W]
    mv[..., :2] = blk
    mv[..., 2] = rng.choice([-1, -2, -4], (T, h // 8 + 1, W // 8 + 1)).repeat(8, 1).repeat(8, 2)[:, :h, :W]
    gt = rng.integers(0, 256, (T, 4 * h, 4 * W), dtype=np.uint8) if with_gt else None
    return driver.Sequence(lr, pm, res, unflt, mv, gt, name="syn%d" % seed)


def _model(dev, variant="O2"):
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8(alignment="dual_att" if variant == "O1" else "mv_dcn")
    m.load_state_dict(G.seeded_weights(variant), strict=True)
    return m.to(dev).eval()


def _naive_loop(model, q, sid, seed, dev):
    """The reference's loop shape: rebuild the whole window on the host for every frame, no feature cache."""
    T, h, W = q.shape
    H = driver.padded_rows(h)
    frames = []
    for i in range(T):
        o = driver.generate_input_index(i, 7, T - 1)
        def win(a, side):
            idx = [driver.side_info_index(j) if side else j for j in o]
            return torch.from_numpy(R.planes_to_unit(a[idx], H))[None, :, None].to(dev)
        mvs = torch.from_numpy(priors_ref.mv2mvs_model_layout(np.pad(q.mvl0[driver.side_info_index(i)], ((0, H - h), (0, 0), (0, 0)))))
        mvs = priors_ref.modify_mv_for_end_frames(i, mvs.numpy().copy(), T)
        noise = driver.noise_for(seed, [sid], i, H, W, dev)
        sr, _ = model(win(q.lr, False), None, torch.from_numpy(mvs).to(dev), win(q.pm, True), win(q.res, True), win(q.unflt, True),
                      None, noise=noise)
        frames.append(R.sr_to_u8(sr.float().cpu().numpy()[0, 0], 4 * h))
    return frames


@pytest.mark.parametrize("graph", [False, True])
def test_driver_matches_naive_loop(cuda_dev, graph):
    """Cached / pipelined / batched driver == window-by-window loop (T = 6 exercises every end-of-sequence MV fix-up and the
    max(1, i) side-information index); h = 28 exercises the row padding (28 -> 32) and the crop of the SR rows."""
    T, h, W = 6, 28, 40
    seqs = [_sequence(1, T, h, W), _sequence(2, T, h, W)]
    model = _model(cuda_dev)
    got = {}
    drv = driver.FrameDriver(model, seed=9, graph=graph)
    res = drv.run(seqs, seq_ids=[5, 6], sink=lambda s, i, img: got.__setitem__((s, i), img))
    assert len(got) == 2 * T and res["frames"] == T
    psnr_ref, ssim_ref = np.zeros(2), np.zeros(2)
    for s, q in enumerate(seqs):
        ref = _naive_loop(model, q, 5 + s, 9, cuda_dev)
        for i in range(T):
            assert got[(s, i)].shape == (4 * h, 4 * W)
            d = np.abs(got[(s, i)].astype(np.int64) - ref[i].astype(np.int64))
            # batch 2 vs batch 1 changes reduction orders / tile schedules: ~1e-4 differences in [0, 1] flip the uint8
            # truncation of a few percent of the pixels by one level (north-star bound: 1e-2 = 2.55 levels)
            assert d.max() <= 1 and (d > 0).mean() < 0.08, (s, i, d.max(), (d > 0).mean())
            psnr_ref[s] += R.psnr_y(got[(s, i)], q.gt[i]) / T
            ssim_ref[s] += R.ssim_y(got[(s, i)], q.gt[i]) / T
    assert np.allclose(res["psnr"], psnr_ref, atol=1e-9) and np.allclose(res["ssim"], ssim_ref, atol=1e-10)


def test_driver_is_batching_invariant(cuda_dev):
    """Noise is keyed by (seed, sequence id, frame, neighbour): a sequence gives the same SR frames alone or in a batch."""
    T, h, W = 4, 24, 40
    a, b = _sequence(3, T, h, W, with_gt=False), _sequence(4, T, h, W, with_gt=False)
    model = _model(cuda_dev, "O1")
    both, alone = {}, {}
    driver.FrameDriver(model, seed=1).run([a, b], seq_ids=[10, 11], sink=lambda s, i, img: both.__setitem__((s, i), img))
    driver.FrameDriver(model, seed=1).run([b], seq_ids=[11], sink=lambda s, i, img: alone.__setitem__((s, i), img))
    for i in range(T):
        d = np.abs(both[(1, i)].astype(np.int64) - alone[(0, i)].astype(np.int64))
        assert d.max() <= 1 and (d > 0).mean() < 0.08


def test_run_sharded_single_process(cuda_dev):
    T, h, W = 3, 24, 24
    seqs = [_sequence(20 + k, T, h, W) for k in range(3)]
    model = _model(cuda_dev)
    r = driver.run_sharded(model, seqs, batch=2, seed=2)
    assert r["frames"] == [3.0, 3.0, 3.0] and len(r["psnr"]) == 3 and all(np.isfinite(r["psnr"])) and all(0 < s < 1 for s in r["ssim"])
    one = driver.FrameDriver(model, seed=2).run([seqs[2]], seq_ids=[2])
    # the feature extraction still runs on cuDNN, whose algorithm choice (and with it the fp32 summation order) can change with
    # the allocator state: ~4e-4 on the SR frame = single uint8 levels on a few pixels
    assert abs(one["psnr"][0] - r["psnr"][2]) < 5e-3 and abs(one["ssim"][0] - r["ssim"][2]) < 1e-4
