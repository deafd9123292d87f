"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the evaluation loop's frame conversions and PSNR / SSIM.

Not imported by the product package (cdfo_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU arm may use it.
Pinned by tests/golden/metrics_golden.npz, which oracle/make_golden_metrics.py produced by calling the reference's own
metric/psnr_ssim.py functions in the build container.

  planes_to_unit   test_LD_37.py:19-29   img.astype(float32) / 255.0, 270-row frames padded with two zero rows (:24-26)
  sr_to_u8         test_LD_37.py:172-180 crop of the padded rows, clamp(0, 1) * 255.0, astype(uint8)
  psnr_y / ssim_y  metric/psnr_ssim.py:278-317, :320-350, :353-399 as called by cal_psnr_ssim :446-484
                   (crop_border = 4, test_y_channel = True -> the float32 round trip of to_y_channel :201-214)
"""
import numpy as np


def planes_to_unit(planes, rows_out=None):
    """[..., H, W] integer planes -> float32 / 255 with zero rows appended up to rows_out."""
    y = planes.astype(np.float32) / np.float32(255.0)
    if rows_out is not None and rows_out > y.shape[-2]:
        pad = np.zeros(y.shape[:-2] + (rows_out - y.shape[-2], y.shape[-1]), np.float32)
        y = np.concatenate([y, pad], axis=-2)
    return y


def sr_to_u8(sr, rows_out):
    """SR [..., H, W] float32 -> uint8 [..., rows_out, W] (test_LD_37.py:172-180: slice, clamp, * 255.0, astype(uint8))."""
    out = np.clip(sr[..., :rows_out, :].astype(np.float32), 0.0, 1.0) * np.float32(255.0)
    return out.astype(np.uint8)


def _to_y(img):
    """to_y_channel for a single-channel image (metric/psnr_ssim.py:210-214)."""
    return img.astype(np.float32) / np.float32(255.0) * np.float32(255.0)


def gaussian_kernel_11():
    """cv2.getGaussianKernel(11, 1.5) (metric/psnr_ssim.py:334)."""
    x = np.arange(11, dtype=np.float64) - 5.0
    k = np.exp(-(x * x) / (2.0 * 1.5 * 1.5))
    return k / k.sum()


def psnr_y(res_u8, gt_u8, border=4):
    a = _to_y(res_u8.astype(np.float64)[border:-border, border:-border])
    b = _to_y(gt_u8.astype(np.float64)[border:-border, border:-border])
    mse = np.mean(((a - b) ** 2).astype(np.float64))   # the reference averages the float32 squares in float32 (<= 1e-6 relative)
    if mse == 0:
        return float("inf")
    return float(20.0 * np.log10(255.0 / np.sqrt(mse)))


def ssim_y(res_u8, gt_u8, border=4):
    a = _to_y(res_u8.astype(np.float64)[border:-border, border:-border]).astype(np.float64)
    b = _to_y(gt_u8.astype(np.float64)[border:-border, border:-border]).astype(np.float64)
    k = gaussian_kernel_11()
    win = np.outer(k, k)
    H, W = a.shape
    Hs, Ws = H - 10, W - 10

    def filt(img):   # cv2.filter2D(img, -1, window)[5:-5, 5:-5] = valid correlation
        out = np.zeros((Hs, Ws), np.float64)
        for i in range(11):
            for j in range(11):
                out += win[i, j] * img[i:i + Hs, j:j + Ws]
        return out

    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    mu1, mu2 = filt(a), filt(b)
    s1 = filt(a * a) - mu1 * mu1
    s2 = filt(b * b) - mu2 * mu2
    s12 = filt(a * b) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return float(m.mean())
