"""Shared by tests: seeded inputs identical to the ones oracle/make_golden.py fed the real reference."""
import json
import os

import numpy as np
import torch

from cdfo_b200 import synthetic
from oracle import priors_ref
from oracle.make_golden import frame_inputs, module_inputs, priors_cases  # noqa: F401  (no reference import inside)
from oracle.make_golden_dsta import dsta_inputs  # noqa: F401
from oracle.make_golden_c3 import c3_frames, c3_target, psnr as psnr_per_sequence  # noqa: F401  (same: seeds -> inputs only)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def state_dict_template(variant):
    shapes = json.load(open(os.path.join(GOLD, "state_dict_%s.json" % variant)))
    return {k: torch.empty(v) for k, v in shapes.items()}


def seeded_weights(variant):
    return synthetic.seeded_state_dict(state_dict_template(variant), seed=4)


def two_frames():
    """(clip0, mvs0, noise0), (clip1, mvs1, noise1): the two 64x64 windows of model_golden.npz."""
    clip0, mvs0 = frame_inputs(1)
    clip1, mvs1 = frame_inputs(2)
    for k in ("x", "pms", "rms", "ufs"):
        clip1[k] = torch.cat([clip0[k][:, 1:], clip1[k][:, -1:]], 1)
    n0 = synthetic.gumbel_uniforms(4, 0, 0, 1, 64, 64)
    n1 = synthetic.gumbel_uniforms(4, 0, 1, 1, 64, 64)
    return (clip0, mvs0, n0), (clip1, mvs1, n1)


def psnr(a, b, border=4, peak=1.0):
    """PSNR as metric/psnr_ssim.py:278-317 computes it on [0,255] images with crop_border=4 (here on [0,1])."""
    a = np.asarray(a, np.float64)[..., border:-border, border:-border]
    b = np.asarray(b, np.float64)[..., border:-border, border:-border]
    mse = np.mean((a * 255.0 - b * 255.0) ** 2)
    return float("inf") if mse == 0 else 20.0 * np.log10(255.0 / np.sqrt(mse))
