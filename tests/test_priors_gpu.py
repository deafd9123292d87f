"""GPU parity of the prior-decoding kernels: bit-exact integer/IEEE work (A1, A2, A3 indices)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as O
from oracle import priors_ref as P

pytestmark = pytest.mark.gpu


def _mv_field(H, W, seed, dtype=np.int8):
    rng = np.random.default_rng(seed)
    mv = rng.integers(-128, 128, (H, W, 3)).astype(dtype)
    mv[..., 2] = rng.choice([-1, -2, -4, 1, 3], (H, W))
    mv[:2, :, 2] = 0          # ref-distance 0: x/0 -> +-inf, 0/0 -> NaN -> 0
    mv[:1, : W // 2, :2] = 0
    return mv


@pytest.mark.parametrize("dtype", [np.int8, np.int32])
def test_mv2mvs_bit_exact(cuda_dev, dtype):
    import cdfo_b200
    for (H, W, seed) in [(16, 24, 0), (1, 1, 1), (120, 208, 2), (272, 480, 3)]:
        mv = _mv_field(H, W, seed, dtype)
        ref = P.mv2mvs_model_layout(mv)
        out = cdfo_b200.mv2mvs(torch.from_numpy(mv).to(cuda_dev)).cpu().numpy()
        assert out.shape == ref.shape == (1, 7, 2, H, W)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))   # bit pattern, incl. -0.0 / inf


def test_modify_mv_for_end_frames_bit_exact(cuda_dev):
    import cdfo_b200
    mv = _mv_field(8, 12, 7)
    base = np.concatenate([P.mv2mvs_model_layout(mv), P.mv2mvs_model_layout(mv[::-1].copy())], 0)
    for i, mx in [(0, 10), (1, 10), (2, 10), (9, 10), (8, 10), (7, 10), (5, 10), (0, 3), (1, 3), (2, 3), (0, 1)]:
        ref = P.modify_mv_for_end_frames(i, base.copy(), mx)
        t = torch.from_numpy(base.copy()).to(cuda_dev)
        out = cdfo_b200.modify_mv_for_end_frames(i, t, mx)
        assert out.data_ptr() == t.data_ptr()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32)), (i, mx)


def test_flow_warp_values_and_indices(cuda_dev):
    import cdfo_b200
    g = torch.Generator().manual_seed(0)
    for (B, C, H, W) in [(2, 5, 16, 24), (1, 64, 64, 64), (1, 3, 120, 208), (1, 1, 1, 1)]:
        x = torch.randn(B, C, H, W, generator=g)
        # flows as the model gets them: multiples of 1/128 from mv2mvs, plus a few far out of frame
        flow = torch.randint(-64 * 3, 64 * 3, (B, H, W, 2), generator=g).float() / 128.0
        flow[:, 0, 0] = 1e4
        flow[:, -1, -1] = -1e4
        ref, ref_idx = O.flow_warp(x.numpy(), flow.numpy(), formula=0, return_index=True)
        y, idx = cdfo_b200.flow_warp(x.to(cuda_dev), flow.to(cuda_dev), return_index=True)
        assert np.array_equal(idx.cpu().numpy(), ref_idx)                  # floor indices: bit-exact
        assert np.abs(y.cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_flow_warp_integer_flow_columns(cuda_dev):
    """The normalise/un-normalise round trip moves floor() at some columns for integer flows;
    the kernel must replay it, not shortcut to w + flow."""
    import cdfo_b200
    for W in (64, 208, 480):
        H = 8
        x = torch.randn(1, 2, H, W)
        flow = torch.zeros(1, H, W, 2)
        flow[..., 0] = 5.0
        _, ref_idx = O.flow_warp(x.numpy(), flow.numpy(), formula=0, return_index=True)
        _, idx = cdfo_b200.flow_warp(x.to(cuda_dev), flow.to(cuda_dev), return_index=True)
        assert np.array_equal(idx.cpu().numpy(), ref_idx)
        assert (ref_idx[0, 0, :, 1] != np.arange(W) + 5).any()


def test_pack_unpack_c8_roundtrip(cuda_dev):
    import cdfo_b200
    L = cdfo_b200._lib
    x = torch.randn(2, 64, 9, 13, device=cuda_dev)
    c8 = torch.empty((2, 8, 9, 13, 8), dtype=torch.bfloat16, device=cuda_dev)
    L.check(L.lib().cdfo_pack_c8(L.ptr(x), L.ptr(c8), 2, 64, 9, 13, L.stream_ptr(cuda_dev)))
    expect = x.view(2, 8, 8, 9, 13).permute(0, 1, 3, 4, 2).to(torch.bfloat16)
    assert torch.equal(c8, expect)
    back = torch.empty_like(x)
    L.check(L.lib().cdfo_unpack_c8(L.ptr(c8), L.ptr(back), 2, 64, 9, 13, L.stream_ptr(cuda_dev)))
    assert torch.equal(back, x.to(torch.bfloat16).float())


def test_prior_conv_vs_torch(cuda_dev):
    """conv_expand_ufs / conv_expand_rms (arch/SIDECVSR_our.py:4446-4447): Conv2d(1, 64, 3, 1, 1), fp32, ragged size."""
    import torch.nn as nn
    import torch.nn.functional as F
    from cdfo_b200 import hotpath
    g = torch.Generator().manual_seed(2)
    conv = nn.Conv2d(1, 64, 3, 1, 1)
    x = torch.randn(3, 1, 19, 37, generator=g)
    with torch.no_grad():
        ref = F.conv2d(x, conv.weight, conv.bias, padding=1)
    got = hotpath.prior_conv(conv.to(cuda_dev), x.to(cuda_dev)).cpu()
    assert (got - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("dtype", [np.int8, np.int32])
def test_mv2mvs_ra_bit_exact(cuda_dev, dtype):
    """RA MV decoding on the device: bit-exact against the golden vector of the reference's Augment and against the
    oracle on a larger random field with sentinels."""
    import cdfo_b200
    import golden_util as G
    from oracle import priors_ref
    g = G.load("priors_ra_golden.npz")
    l0, l1 = g["l0"][0].astype(dtype), g["l1"][0].astype(dtype)
    got = cdfo_b200.mv2mvs_ra(torch.from_numpy(l0).to(cuda_dev), torch.from_numpy(l1).to(cuda_dev)).cpu().numpy()
    ref = np.ascontiguousarray(g["flows"][None].transpose(0, 1, 4, 2, 3))
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    rng = np.random.default_rng(3)
    H, W = 40, 56
    a = rng.integers(-128, 128, (H, W, 3)).astype(dtype)
    b = rng.integers(-128, 128, (H, W, 3)).astype(dtype)
    a[..., 2] = rng.choice([-1, -2, -4, -99, 0], (H, W))
    b[..., 2] = rng.choice([1, 2, 4, -99, 0], (H, W))
    got = cdfo_b200.mv2mvs_ra(torch.from_numpy(a).to(cuda_dev), torch.from_numpy(b).to(cuda_dev)).cpu().numpy()
    ref = np.ascontiguousarray(priors_ref.mv2mvs_ra(a, b)[None].transpose(0, 1, 4, 2, 3))
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
