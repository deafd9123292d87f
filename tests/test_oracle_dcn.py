"""Pins the plain-C oracle (oracle/dcn_ref.c) before anything trusts it:
the reference's only known-answer vector (ops/dcn/simple_check.py:11-22) and
torchvision's CPU deform_conv2d, the op the model's DCN alignment calls
(arch/SIDECVSR_our.py:3352)."""
import numpy as np
import pytest
import torch
import torchvision

from oracle import c_oracle as O


def simple_check_case():
    """ops/dcn/simple_check.py: DeformConv(2,1,3,padding=1,deformable_groups=2), weight == 1."""
    off = np.array([1, 1, 1, 0, 1, -1, 0, 1, 0, 0, 0, -1, -1, 1, -1, 0, -1, -1], np.float32)
    off = np.tile(off[None, :, None, None], (1, 2, 3, 3))
    x = np.arange(18, dtype=np.float32).reshape(1, 2, 3, 3)
    w = np.ones((1, 2, 3, 3), np.float32)
    gt = np.array([81, 99, 117, 135, 153, 171, 189, 207, 225], np.float32)
    return x, off, w, gt


def test_simple_check_known_answer():
    x, off, w, gt = simple_check_case()
    y = O.dcn_forward(x, off, None, w, None, stride=1, padding=1, dilation=1, groups=1, deformable_groups=2)
    assert np.abs(gt - y.flatten()).sum() < 1e-8


def test_simple_check_torchvision_agrees():
    x, off, w, gt = simple_check_case()
    y = torchvision.ops.deform_conv2d(torch.from_numpy(x), torch.from_numpy(off), torch.from_numpy(w), None, 1, 1, 1)
    assert np.abs(gt - y.numpy().flatten()).sum() < 1e-8


CASES = [
    # B, C, H, W, Co, k, stride, pad, dil, groups, dg, use_mask, use_bias
    (2, 8, 9, 11, 6, 3, 1, 1, 1, 1, 4, True, True),
    (1, 8, 9, 11, 6, 3, 2, 2, 2, 2, 4, True, True),
    (1, 4, 7, 5, 4, 1, 1, 0, 1, 1, 1, True, False),
    (2, 6, 8, 8, 3, 3, 1, 1, 1, 3, 2, False, False),
    (1, 64, 12, 16, 64, 3, 1, 1, 1, 1, 16, True, True),   # the model's hot shape (small H, W)
    (1, 16, 10, 13, 16, 3, 1, 1, 1, 1, 16, True, True),   # DSTA's internal DCN shape (dg == C)
]


@pytest.mark.parametrize("case", CASES)
def test_oracle_vs_torchvision_cpu(case):
    B, C, H, W, Co, k, s, p, d, groups, dg, use_mask, use_bias = case
    g = torch.Generator().manual_seed(hash(case) % (2**31))
    Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1
    Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    x = torch.randn(B, C, H, W, generator=g)
    offset = torch.randn(B, dg * 2 * k * k, Ho, Wo, generator=g) * 3.0
    mask = torch.rand(B, dg * k * k, Ho, Wo, generator=g) if use_mask else None
    wt = torch.randn(Co, C // groups, k, k, generator=g) * 0.2
    b = torch.randn(Co, generator=g) if use_bias else None
    ref = torchvision.ops.deform_conv2d(x, offset, wt, b, s, p, d, mask).numpy()
    y = O.dcn_forward(x.numpy(), offset.numpy(), None if mask is None else mask.numpy(), wt.numpy(),
                      None if b is None else b.numpy(), s, p, d, groups, dg)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_border_cases_known_answers():
    """Hand-computed border behaviour: a sample is dropped only when h<=-1, w<=-1, h>=H or w>=W; corners that
    fall outside contribute zero (deform_conv_cuda_kernel.cu:617, :480-490)."""
    x = np.ones((1, 1, 4, 4), np.float32)
    w = np.ones((1, 1, 1, 1), np.float32)
    m = np.ones((1, 1, 4, 4), np.float32)

    def run(dy, dx):
        off = np.zeros((1, 2, 4, 4), np.float32)
        off[:, 0], off[:, 1] = dy, dx
        return O.dcn_forward(x, off, m, w, None, 1, 0, 1, 1, 1)[0, 0]

    y = run(-0.5, 0.0)   # row 0 samples h=-0.5: half of the weight falls on the zero border
    assert np.allclose(y[0], 0.5) and np.allclose(y[1:], 1.0)
    y = run(-1.0, 0.0)   # row 0 samples h=-1 exactly: dropped by the inside test
    assert np.allclose(y[0], 0.0) and np.allclose(y[1:], 1.0)
    y = run(0.0, 0.75)   # last column samples w=3.75: 0.25 of the weight inside
    assert np.allclose(y[:, 3], 0.25) and np.allclose(y[:, :3], 1.0)
    y = run(0.0, 1.0)    # last column samples w=4 == W: dropped
    assert np.allclose(y[:, 3], 0.0)


def test_shape_errors():
    x = np.zeros((1, 6, 4, 4), np.float32)
    with pytest.raises(ValueError):
        O.dcn_forward(x, np.zeros((1, 72, 4, 4), np.float32), None, np.zeros((4, 6, 3, 3), np.float32), None,
                      1, 1, 1, 1, 4)  # C % dg != 0


def test_flow_warp_oracle_index_formulas():
    """Integer-valued flows: floor(round trip) differs from floor(w + flow) at some columns;
    the two ATen un-normalise formulas are both restated and stay within 1 ulp-level value agreement."""
    H, W = 8, 208
    x = np.random.default_rng(0).standard_normal((1, 2, H, W)).astype(np.float32)
    flow = np.zeros((1, H, W, 2), np.float32)
    flow[..., 0] = 3.0
    y0, i0 = O.flow_warp(x, flow, formula=0, return_index=True)
    y1, i1 = O.flow_warp(x, flow, formula=1, return_index=True)
    naive = np.arange(W) + 3
    assert (i0[0, 0, :, 1] != naive).sum() > 0          # the round trip does move some indices
    assert np.abs(y0 - y1).max() < 1e-4
    expect = np.zeros_like(x)
    expect[..., : W - 3] = x[..., 3:]
    assert np.abs(y0 - expect).max() < 1e-3
