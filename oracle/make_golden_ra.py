"""TEST INFRASTRUCTURE ONLY -- golden vectors for the RA twin of the MV decoding.

Run in the build container (needs /root/reference).  Imports the reference's own `opt/data_RA_bi.py::Augment`
(train-time decoding of an (l0, l1) MV pair into seven flows, data_RA_bi.py:419-424, :496-533) with `skimage` stubbed
(imported at the top of that file, unused by Augment), calls it with hflip=False, rot=False (no random augmentation)
and applies the train loop's `/ 32.0` (train_RA_37.py:383-386).  Output: tests/golden/priors_ra_golden.npz.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference"


def ra_case():
    """int8 l0 / l1 fields [1,H,W,3] = (mv_a, mv_b, refdist): l0 refdist < 0, l1 > 0, -99 sentinels in either list, zeros."""
    rng = np.random.default_rng(11)
    H, W = 16, 24
    l0 = rng.integers(-128, 128, (1, H, W, 3)).astype(np.int8)
    l1 = rng.integers(-128, 128, (1, H, W, 3)).astype(np.int8)
    l0[..., 2] = rng.choice([-1, -2, -4, -8], (1, H, W))
    l1[..., 2] = rng.choice([1, 2, 4, 8], (1, H, W))
    l0[0, 2:5, :, 2] = -99                       # l0 missing -> complemented from l1
    l1[0, 4:7, :, 2] = -99                       # l1 missing (rows 4: both missing)
    l0[0, 8, :6, :2] = 0                         # zero vectors
    l1[0, 9, :6, 2] = 0                          # x / 0 = +-inf, 0 / 0 = nan -> 0
    l1[0, 9, :3, :2] = 0
    return l0, l1


def main():
    sys.path.insert(0, REF)
    sk = types.ModuleType("skimage")
    sk.io, sk.transform = types.ModuleType("skimage.io"), types.ModuleType("skimage.transform")
    sys.modules.update({"skimage": sk, "skimage.io": sk.io, "skimage.transform": sk.transform})
    import opt.data_RA_bi as D
    l0, l1 = ra_case()
    H, W = l0.shape[1:3]
    z = np.zeros((7, H, W), np.float32)
    sample = {"lr_imgs": z, "hr_imgs": np.zeros((7, 4 * H, 4 * W), np.float32), "mvl0s": l0.copy(), "mvl1s": l1.copy(),
              "res_s": z, "mpm_s": z, "pred_fs": z, "unflt_fs": z, "qp": 37, "lrbi": z}
    with np.errstate(all="ignore"):
        out = D.Augment()(sample, hflip=False, rot=False)
    flows = np.asarray(out["mvl0s"], np.float32) / np.float32(32.0)      # train_RA_37.py:384
    assert out["mvl1s"] is not None and np.array_equal(np.asarray(out["mvl1s"]), np.asarray(out["mvl0s"]), equal_nan=True)
    np.savez_compressed(os.path.join(GOLD, "priors_ra_golden.npz"), l0=l0, l1=l1, flows=flows)
    print("wrote priors_ra_golden.npz", flows.shape, float(np.nanmax(np.abs(flows[np.isfinite(flows)]))))


if __name__ == "__main__":
    main()
