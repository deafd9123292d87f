"""TEST INFRASTRUCTURE ONLY -- golden vectors for the PSNR / SSIM of the evaluation loop.

Run in the build container (needs /root/reference): imports the reference's own metric/psnr_ssim.py and calls
calculate_psnr / calculate_ssim exactly like cal_psnr_ssim does (metric/psnr_ssim.py:462-472: [H,W,1] float64 arrays,
crop_border 4, test_y_channel True).  Output: tests/golden/metrics_golden.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def cases():
    rng = np.random.default_rng(5)
    out = []
    for (H, W, noise) in ((40, 52, 6), (64, 96, 20), (33, 47, 2), (72, 64, 60)):
        yy, xx = np.mgrid[0:H, 0:W]
        base = 128 + 90 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + rng.normal(0, 12, (H, W))
        gt = np.clip(np.round(base), 0, 255).astype(np.uint8)
        res = np.clip(gt.astype(np.int64) + rng.integers(-noise, noise + 1, gt.shape), 0, 255).astype(np.uint8)
        out.append((res, gt))
    flat = np.full((30, 30), 77, np.uint8)
    out.append((flat.copy(), flat.copy()))          # identical images: PSNR = inf, SSIM = 1
    return out


def main():
    sys.path.insert(0, "/root/reference")
    from metric import psnr_ssim as M
    data = {}
    for i, (res, gt) in enumerate(cases()):
        a, b = res[:, :, None].astype(np.float64), gt[:, :, None].astype(np.float64)
        data["res%d" % i], data["gt%d" % i] = res, gt
        data["psnr%d" % i] = np.float64(M.calculate_psnr(a, b, 4, test_y_channel=True))
        data["ssim%d" % i] = np.float64(M.calculate_ssim(a, b, 4, test_y_channel=True))
        print(i, res.shape, data["psnr%d" % i], data["ssim%d" % i])
    np.savez_compressed(os.path.join(GOLD, "metrics_golden.npz"), n=np.int64(len(cases())), **data)


if __name__ == "__main__":
    main()
