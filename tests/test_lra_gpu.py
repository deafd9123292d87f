"""Prior-guided long-range attention kernels (csrc/lra.cu) against the torch oracle (pinned to the reference)."""
import numpy as np
import pytest
import torch

import golden_util as G
from oracle import torch_ref

pytestmark = pytest.mark.gpu


def _model(dev):
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8()
    m.load_state_dict(G.seeded_weights("O1"), strict=True)
    return m.to(dev).eval()


def _inputs(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 64, H, W, generator=g) * 0.3, torch.randn(B, 64, H, W, generator=g) * 0.5,
            torch.rand(B, 64, H, W, generator=g).clamp_min(1e-12))


@pytest.mark.parametrize("B,H,W", [(1, 8, 8), (2, 24, 40), (1, 40, 24), (1, 64, 64), (1, 120, 208)])
def test_lra_vs_oracle(cuda_dev, B, H, W):
    res, x, u = _inputs(B, H, W, seed=H * 1000 + W)
    sd = G.seeded_weights("O1")
    with torch.no_grad():
        ref = torch_ref.long_range_attention(sd, "RDAB.", res, x, u)
        mask = torch_ref.lra_mask(sd, "RDAB.", res, u)
    m = _model(cuda_dev)
    out = m.RDAB(res.to(cuda_dev), x.to(cuda_dev), u.to(cuda_dev)).cpu()
    err = (out - ref).abs().max().item()
    frac = mask.sum(1).clamp(max=1).mean().item()
    print("LRA B%d %dx%d: max err %.3g, max|ref| %.3g, masked tokens %.1f%%" % (B, H, W, err, ref.abs().max().item(), 100 * frac))
    assert mask.sum(1).max().item() <= 1          # the structural fact the kernels rely on
    assert err <= 2e-3


def test_lra_dense_mask(cuda_dev):
    """Every token masked on the same / neighbouring channels (a trained model can do this): exercises the
    per-query softmax path and the R[c1][c2] overlap table at every (w, w') pair."""
    B, H, W = 1, 16, 32
    res, x, u = _inputs(B, H, W, seed=5)
    sd = dict(G.seeded_weights("O1"))
    sd["RDAB.conv_du_re2.0.bias"] = sd["RDAB.conv_du_re2.0.bias"].clone()
    sd["RDAB.conv_du_re2.0.bias"][10] += 30.0      # channel 10 wins the softmax everywhere
    u = u.clone()
    u[:, 13, :, ::3] = 1.0 - 1e-7                  # ... except every third column, where channel 13's Gumbel noise is huge
    sd["RDAB.conv_du_re2.0.bias"][13] += 16.0
    with torch.no_grad():
        ref = torch_ref.long_range_attention(sd, "RDAB.", res, x, u)
        mask = torch_ref.lra_mask(sd, "RDAB.", res, u)
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8()
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda_dev).eval()
    out = m.RDAB(res.to(cuda_dev), x.to(cuda_dev), u.to(cuda_dev)).cpu()
    frac = mask.sum(1).clamp(max=1).mean().item()
    chans = sorted(set(mask.sum(dim=(0, 2, 3)).nonzero().flatten().tolist()))
    err = (out - ref).abs().max().item()
    print("LRA dense mask: masked tokens %.1f%% on channels %s, max err %.3g" % (100 * frac, chans, err))
    assert frac > 0.9 and len(chans) >= 2
    assert err <= 2e-3


def test_lra_c8_output_is_the_rounded_fp32_output(cuda_dev):
    """long_range_attention(out8=...): the fuse kernel packs its result as bf16 into a channel range of the consumer's c8 tensor
    (the model's cat([fea, x_n]), arch:4454); equals the bf16 rounding of the fp32 output, other channels untouched."""
    from cdfo_b200 import conv, hotpath
    B, H, W = 2, 24, 40
    res, x, u = _inputs(B, H, W, seed=77)
    m = _model(cuda_dev)
    res, x, u = res.to(cuda_dev), x.to(cuda_dev), u.to(cuda_dev)
    ref = hotpath.long_range_attention(m.RDAB, res, x, u, x2=res)
    out8 = torch.full((B, 16, H, W, 8), 3.0, dtype=torch.bfloat16, device=cuda_dev)
    hotpath.long_range_attention(m.RDAB, res, x, u, x2=res, out8=out8, channel0=64)
    conv.to_c8(x, out=out8, channel0=0)
    got = conv.from_c8(out8)
    assert torch.equal(got[:, 64:], ref.to(torch.bfloat16).float())
    assert torch.equal(got[:, :64], x.to(torch.bfloat16).float())


@pytest.mark.parametrize("K,Co,mode,H,W", [(64, 64, 0, 9, 7), (64, 128, 0, 10, 13), (128, 64, 1, 5, 9), (64, 128, 0, 16, 20), (128, 64, 1, 12, 12)])
def test_pointwise_conv_shapes(cuda_dev, K, Co, mode, H, W):
    """Tensor-core 1x1 convolution (TF32): pipelined kernel (H*W % 4 == 0, ragged last tile) and the scalar fallback (odd sizes)."""
    from cdfo_b200 import hotpath
    g = torch.Generator().manual_seed(K + Co + H)
    B = 2
    wgt = (torch.randn(Co, K, generator=g) / K ** 0.5).to(cuda_dev)
    bias = torch.randn(Co, generator=g).to(cuda_dev)
    r1 = torch.randn(B, Co, H, W, generator=g).to(cuda_dev)
    if mode == 0:
        a, b2 = torch.randn(B, K, H, W, generator=g).to(cuda_dev), torch.randn(B, K, H, W, generator=g).to(cuda_dev)
        ref = torch.einsum("ok,bkhw->bohw", wgt.double(), (a + b2).double()) + bias.view(1, -1, 1, 1).double()
    else:
        a, b2 = torch.randn(B, H * W, 64, generator=g).to(cuda_dev), torch.randn(B, H * W, 64, generator=g).to(cuda_dev)
        cat = torch.cat([a, b2], 2).double()
        ref = torch.einsum("ok,bpk->bop", wgt.double(), cat).view(B, Co, H, W) + bias.view(1, -1, 1, 1).double()
    ref = torch.relu(ref) + r1.double()
    out = torch.empty((B, Co, H, W), dtype=torch.float32, device=cuda_dev)
    from cdfo_b200 import _lib
    _lib.call("cdfo_pointwise_conv_fwd", _lib.ptr(a), _lib.ptr(b2), _lib.ptr(wgt), _lib.ptr(bias), _lib.ptr(r1), _lib.ptr(None), _lib.ptr(out),
              B, K, Co, H, W, 1, mode, _lib.stream_ptr(cuda_dev))
    err = (out.double() - ref).abs().max().item()
    assert err <= 5e-3 * max(1.0, ref.abs().max().item()), err


def test_lra_col_bf16_vs_tf32(cuda_dev):
    """The bf16 mma.sync column pass against the TF32 one: same result to bf16-operand accuracy."""
    from cdfo_b200 import _lib
    B, H, W = 2, 72, 40          # H not a multiple of 64: masked keys in the last block; 5 query tiles over 9 warps
    res, x, u = _inputs(B, H, W, seed=3)
    m = _model(cuda_dev)
    res, x, u = res.to(cuda_dev), x.to(cuda_dev), u.to(cuda_dev)
    try:
        _lib.call("cdfo_lra_set_col_tcgen05", 0)
        _lib.call("cdfo_lra_set_col_precision", 1)
        ref = m.RDAB(res, x, u)
        _lib.call("cdfo_lra_set_col_precision", 0)
        out = m.RDAB(res, x, u)
    finally:
        _lib.call("cdfo_lra_set_col_precision", 0)
        _lib.call("cdfo_lra_set_col_tcgen05", 1)
    err = (out - ref).abs().max().item()
    print("LRA column pass bf16 vs tf32: max diff %.3g (max|out| %.3g)" % (err, ref.abs().max().item()))
    assert err <= 4e-3


@pytest.mark.parametrize("B,H,W", [(2, 72, 40), (1, 136, 24), (2, 272, 16), (1, 40, 8), (1, 264, 8)])
def test_lra_col_tcgen05_vs_tf32(cuda_dev, B, H, W):
    """The tcgen05 column pass (csrc/lra_col_sm100.cu: score tile in tensor memory, the default) against the TF32 mma.sync pass:
    one / two / three 128-query tiles, one and two key chunks, H not a multiple of 16 * 8, rerun bit-identical."""
    from cdfo_b200 import _lib
    res, x, u = _inputs(B, H, W, seed=H + W)
    m = _model(cuda_dev)
    res, x, u = res.to(cuda_dev), x.to(cuda_dev), u.to(cuda_dev)
    try:
        _lib.call("cdfo_lra_set_col_precision", 1)
        ref = m.RDAB(res, x, u)
    finally:
        _lib.call("cdfo_lra_set_col_precision", 0)
    out = m.RDAB(res, x, u)
    err = (out - ref).abs().max().item()
    print("LRA column pass tcgen05 vs tf32 %dx%dx%d: max diff %.3g (max|out| %.3g)" % (B, H, W, err, ref.abs().max().item()))
    assert err <= 4e-3
    assert torch.equal(m.RDAB(res, x, u), out)


@pytest.mark.parametrize("B,H,W", [(2, 24, 40), (1, 30, 34), (3, 64, 72), (1, 272, 480)])
def test_mask_logits_kernel_vs_torch(cuda_dev, B, H, W):
    """csrc/lra_mask_logits.cu (stride-2 convolution on TF32 tensor cores + ReLU + mean + 1x1 + ReLU, arch:2183-2186) against fp32 torch."""
    import torch.nn.functional as F
    from cdfo_b200 import hotpath
    from cdfo_b200.model import LLongRangAttention
    torch.manual_seed(B * H + W)
    mod = LLongRangAttention(64).to(cuda_dev)
    v = torch.relu(torch.randn(B, 64, H, W) * 0.5)
    c2, c3 = mod.conv_du_re._modules["2"], mod.conv_du_re2._modules["0"]
    with torch.no_grad():
        r = F.relu(F.conv2d(v.double(), c2.weight.detach().cpu().double(), c2.bias.detach().cpu().double(), stride=2, padding=2))
        ref = F.relu(F.conv2d(r.mean(dim=(2, 3), keepdim=True), c3.weight.detach().cpu().double(), c3.bias.detach().cpu().double())).reshape(B, 64)
    got = hotpath.mask_logits(mod, v.to(cuda_dev)).cpu().double()
    err = (got - ref).abs().max().item()
    print("mask logits %dx%dx%d: max err %.3g (max|ref| %.3g)" % (B, H, W, err, ref.abs().max().item()))
    assert err <= 2e-4 * max(1.0, ref.abs().max().item())
    assert torch.equal(hotpath.mask_logits(mod, v.to(cuda_dev)).cpu().double(), got)      # fixed-order reduction


@pytest.mark.parametrize("B,H,W,scale", [(2, 24, 40, 0.5), (1, 64, 64, 1.5)])
def test_lra_window_tensor_core_vs_fp32(cuda_dev, B, H, W, scale):
    """The tensor-core window pass (bf16 hi/lo split scores, the default) against round 1's fp32 SIMT kernel, also with logits of
    several tens (scale 1.5: |q|^2 ~ 100), where an unsplit bf16 product would be visibly wrong."""
    from cdfo_b200 import _lib
    res, x, u = _inputs(B, H, W, seed=7 * H + W)
    x = x * scale / 0.5
    m = _model(cuda_dev)
    res, x, u = res.to(cuda_dev), x.to(cuda_dev), u.to(cuda_dev)
    try:
        _lib.call("cdfo_lra_set_win_tensor_core", 0)
        ref = m.RDAB(res, x, u)
    finally:
        _lib.call("cdfo_lra_set_win_tensor_core", 1)
    out = m.RDAB(res, x, u)
    err = (out - ref).abs().max().item()
    print("LRA window pass tensor cores vs fp32 (x scale %.1f): max diff %.3g (max|out| %.3g)" % (scale, err, ref.abs().max().item()))
    assert err <= 4e-3 * max(1.0, ref.abs().max().item())


def test_mask_logits_tma_vs_cp_async(cuda_dev):
    """The TMA-fed stride-2 convolution of the mask logits (one box per stage, mbarrier ring; the default when W % 4 == 0) against
    the cp.async kernel: same TF32 products, different summation order over the input-channel chunks."""
    from cdfo_b200 import _lib, hotpath
    m = _model(cuda_dev)
    g = torch.Generator().manual_seed(11)
    v = torch.rand(3, 64, 46, 52, generator=g).to(cuda_dev)       # ragged tiles in both directions
    try:
        _lib.call("cdfo_lra_set_logit_tma", 0)
        ref = hotpath.mask_logits(m.RDAB, v)
    finally:
        _lib.call("cdfo_lra_set_logit_tma", 1)
    got = hotpath.mask_logits(m.RDAB, v)
    err = (got - ref).abs().max().item()
    print("mask logits TMA vs cp.async: max diff %.3g (max|ref| %.3g)" % (err, ref.abs().max().item()))
    assert err <= 1e-5 * max(1.0, ref.abs().max().item())
    assert torch.equal(hotpath.mask_logits(m.RDAB, v), got)


@pytest.mark.parametrize("B,H,W", [(2, 24, 40), (1, 16, 72), (2, 272, 480)])
def test_lra_row_tma_equals_cp_async(cuda_dev, B, H, W):
    """The row pass fed by tiled TMA (64-pixel pieces through an mbarrier ring, the default) against the 4-byte cp.async load phase:
    same fp32 operations in the same order, so the module output is bit-identical (one, two and eight pieces; a partial last piece)."""
    from cdfo_b200 import _lib
    res, x, u = _inputs(B, H, W, seed=3 * H + W)
    m = _model(cuda_dev)
    res, x, u = res.to(cuda_dev), x.to(cuda_dev), u.to(cuda_dev)
    try:
        _lib.call("cdfo_lra_set_row_tma", 0)
        ref = m.RDAB(res, x, u)
    finally:
        _lib.call("cdfo_lra_set_row_tma", 1)
    out = m.RDAB(res, x, u)
    assert torch.equal(out, ref)


def test_lra_mask_logits_from_one_channel_prior(cuda_dev):
    """res = conv_expand_rms(rms) (arch:4447) feeds conv_du_re.0 + ReLU (arch:2183): composed into ONE direct 1 -> 64 convolution of the
    one-channel map (long_range_attention(res_prior=...)).  The module output must agree with the 64-channel route and with the oracle."""
    from cdfo_b200 import hotpath
    B, H, W = 2, 40, 56
    m = _model(cuda_dev)
    g = torch.Generator().manual_seed(21)
    rms1 = torch.rand(B, 1, H, W, generator=g)
    _, x, u = _inputs(B, H, W, seed=9)
    sd = G.seeded_weights("O1")
    with torch.no_grad():
        res = torch.nn.functional.conv2d(rms1, sd["conv_expand_rms.weight"], sd["conv_expand_rms.bias"], padding=1)
        ref = torch_ref.long_range_attention(sd, "RDAB.", res, x, u)
    d = lambda t: t.to(cuda_dev)
    res_d = hotpath.prior_conv(m.conv_expand_rms, d(rms1))
    a = hotpath.long_range_attention(m.RDAB, res_d, d(x), d(u))
    b = hotpath.long_range_attention(m.RDAB, res_d, d(x), d(u), res_prior=(m.conv_expand_rms, d(rms1)))
    err_a, err_b = (a.cpu() - ref).abs().max().item(), (b.cpu() - ref).abs().max().item()
    print("LRA logits from the one-channel prior: max err %.3g (64-channel route %.3g), routes differ by %.3g"
          % (err_b, err_a, (a - b).abs().max().item()))
    assert err_b <= 2e-3 and err_a <= 2e-3
