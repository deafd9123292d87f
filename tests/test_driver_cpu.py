"""Host logic of the frame driver (cdfo_b200/driver.py) and the metrics oracle against the reference's own numbers.
No GPU: window indices, side-information indexing, the on-disk reader, error behaviour."""
import os

import numpy as np
import pytest

from cdfo_b200 import driver
from oracle import metrics_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_metrics_oracle_matches_reference_golden():
    """oracle/metrics_ref.py vs outputs of metric/psnr_ssim.py (oracle/make_golden_metrics.py)."""
    g = np.load(os.path.join(GOLD, "metrics_golden.npz"))
    for i in range(int(g["n"])):
        res, gt = g["res%d" % i], g["gt%d" % i]
        p, s = R.psnr_y(res, gt), R.ssim_y(res, gt)
        if np.isinf(g["psnr%d" % i]):
            assert np.isinf(p) and s == 1.0
        else:
            assert abs(p - float(g["psnr%d" % i])) < 2e-5     # the reference returns PSNR as float32
            assert abs(s - float(g["ssim%d" % i])) < 1e-12


def test_window_indices_match_reference_formula():
    for T in (1, 2, 3, 5, 7, 12):
        for i in range(T):
            ref = np.clip(np.array(range(7)) - 3 + i, 0, T - 1).tolist()      # test_LD_37.py:13-16
            assert driver.generate_input_index(i, 7, T - 1) == ref
            if i >= 1:   # the cached window = previous window shifted by one + the new frame
                prev = driver.generate_input_index(i - 1, 7, T - 1)
                assert prev[1:] + [driver.new_frame_of_step(i, T)] == ref


def test_side_info_index_quirk():
    assert [driver.side_info_index(j) for j in (0, 1, 2, 9)] == [1, 1, 2, 9]    # "%05d" % max(1, i), test_LD_37.py:36


def test_padded_rows():
    assert driver.padded_rows(270) == 272 and driver.padded_rows(272) == 272      # test_LD_37.py:24-26
    assert driver.padded_rows(180) == 184 and driver.padded_rows(64) == 64        # 736 = 4 * 184 rows at :174-175


def _arrays(T=4, h=6, W=8, seed=0):
    rng = np.random.default_rng(seed)
    return dict(lr=rng.integers(0, 256, (T, h, W), dtype=np.uint8), pm=rng.integers(0, 256, (T, h, W), dtype=np.uint8),
                res=rng.integers(-128, 128, (T, h, W)).astype(np.int64), unflt=rng.integers(0, 256, (T, h, W), dtype=np.uint8),
                mvl0=rng.integers(-64, 64, (T, h, W, 3)).astype(np.int64), gt=rng.integers(0, 256, (T, 4 * h, 4 * W), dtype=np.uint8))


def test_sequence_validation():
    a = _arrays()
    q = driver.Sequence(**a)
    assert q.shape == (4, 6, 8) and q.mvl0.dtype == np.int8 and q.res.dtype == np.int16
    big = dict(a, mvl0=a["mvl0"] * 100)
    assert driver.Sequence(**big).mvl0.dtype == np.int32
    with pytest.raises(ValueError):
        driver.Sequence(**dict(a, pm=a["pm"][:, :5]))
    with pytest.raises(ValueError):
        driver.Sequence(**dict(a, gt=a["gt"][:, :-1]))
    with pytest.raises(ValueError):
        driver.Sequence(**dict(a, mvl0=a["mvl0"].astype(np.float32)))


def test_from_directory_reads_reference_layout(tmp_path):
    cv2 = pytest.importorskip("cv2")
    a = _arrays(T=3, h=8, W=16, seed=3)
    lr_dir, side, gt_dir = tmp_path / "lr" / "Seq_480x272_3F.yuv", tmp_path / "side" / "Seq_480x272_3F", tmp_path / "gt"
    for d in (lr_dir, side / "part_m", side / "res", side / "unfiltered", side / "mvl0", gt_dir):
        os.makedirs(d)
    for t in range(3):
        cv2.imwrite(str(lr_dir / ("%05d.png" % t)), a["lr"][t])
        cv2.imwrite(str(gt_dir / ("%05d.png" % t)), a["gt"][t])
        if t >= 1:   # frame 0 has no side information on disk
            cv2.imwrite(str(side / "part_m" / ("%05d_M_mask.png" % t)), a["pm"][t])
            cv2.imwrite(str(side / "unfiltered" / ("%05d_unflt.png" % t)), a["unflt"][t])
            np.save(str(side / "res" / ("%05d_res.npy" % t)), np.stack([a["res"][t]] * 2, -1))
            np.save(str(side / "mvl0" / ("%05d_mvl0.npy" % t)), a["mvl0"][t])
    q = driver.Sequence.from_directory(str(lr_dir), str(side), str(gt_dir))
    assert q.name == "Seq_480x272_3F.yuv" and q.shape == (3, 8, 16)
    assert np.array_equal(q.lr, a["lr"]) and np.array_equal(q.gt, a["gt"])
    for k in ("pm", "unflt", "res", "mvl0"):
        got = getattr(q, k)
        assert np.array_equal(got[1:], a[k][1:]) and np.array_equal(got[0], a[k][1])   # slot 0 repeats frame 1


def test_driver_is_cuda_only():
    import torch
    m = torch.nn.Linear(1, 1)
    with pytest.raises(NotImplementedError):
        driver.FrameDriver(m)


def test_conversion_oracles():
    x = np.array([[0, 1, 127, 255]], np.uint8)
    y = R.planes_to_unit(x, rows_out=3)
    assert y.shape == (3, 4) and y[0, 3] == 1.0 and y[1:].sum() == 0 and y[0, 1] == np.float32(1) / np.float32(255)
    sr = np.array([[-0.5, 0.0, 0.999, 1.0, 7.0, 0.5]], np.float32)
    assert R.sr_to_u8(sr, 1).tolist() == [[0, 0, 254, 255, 255, 127]]               # truncation, not rounding


def test_feature_ring_window_order_and_wraparound():
    """model.FeatureRing (host logic, runs on CPU tensors): after k pushes the window is the last N frames in temporal order,
    contiguous, with the centre frame and the two neighbour runs as views, for more pushes than there are slots."""
    import torch
    from cdfo_b200.model import FeatureRing
    B, N, C, H, W = 2, 7, 3, 4, 5
    frames = [torch.full((B, C, H, W), float(t)) + torch.arange(B).view(B, 1, 1, 1) * 0.5 for t in range(N + 11)]
    l1 = torch.stack(frames[:N], 1).reshape(B * N, C, H, W)            # the reference's b-major [B*N, C, H, W]
    ring = FeatureRing(l1, B, N)
    handle = ring.handle()
    for t in range(N, N + 11):
        win = ring.window()
        assert win.is_contiguous() and tuple(win.shape) == (N, B, C, H, W)
        for n in range(N):
            assert torch.equal(win[n], frames[t - N + n])
        assert handle._cdfo_ring[1] == ring.serial
        ring.push(frames[t])
        assert handle._cdfo_ring[1] != ring.serial                        # the old handle is stale now
        handle = ring.handle()
        assert handle.data_ptr() == ring.window().data_ptr() and tuple(handle.shape) == (N * B, C, H, W)
    assert torch.equal(ring.window()[N - 1], frames[N + 10])
