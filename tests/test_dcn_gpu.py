"""GPU parity of the deformable-convolution boundary (through the C ABI) against the pinned C oracle."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def exact_fp32_path():
    """These tests pin the fp32 catch-all kernel; the tcgen05 route has its own file (test_sm100_gpu.py)."""
    import cdfo_b200
    saved = cdfo_b200.config.tensor_core
    cdfo_b200.config.tensor_core = False
    yield
    cdfo_b200.config.tensor_core = saved


def _rand_case(case, seed=0):
    B, C, H, W, Co, k, s, p, d, groups, dg, use_mask, use_bias = case
    g = torch.Generator().manual_seed(seed)
    Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1
    Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    x = torch.randn(B, C, H, W, generator=g)
    offset = torch.randn(B, dg * 2 * k * k, Ho, Wo, generator=g) * 3.0
    mask = torch.rand(B, dg * k * k, Ho, Wo, generator=g) if use_mask else None
    wt = torch.randn(Co, C // groups, k, k, generator=g) * 0.2
    b = torch.randn(Co, generator=g) if use_bias else None
    return x, offset, mask, wt, b


def test_simple_check_through_module(cuda_dev):
    """ops/dcn/simple_check.py verbatim in meaning, through the mirrored DeformConv module: exact."""
    import cdfo_b200
    m = cdfo_b200.DeformConv(2, 1, kernel_size=3, padding=1, deformable_groups=2).to(cuda_dev)
    torch.nn.init.constant_(m.weight, 1)
    offset = torch.tensor([1, 1, 1, 0, 1, -1, 0, 1, 0, 0, 0, -1, -1, 1, -1, 0, -1, -1], dtype=torch.float32,
                          device=cuda_dev)
    offset = offset.unsqueeze(0).unsqueeze(-1).unsqueeze(-1).repeat(1, 2, 3, 3)
    inp = torch.arange(18, dtype=torch.float32).view(1, 2, 3, 3).to(cuda_dev)
    gt = torch.tensor([81, 99, 117, 135, 153, 171, 189, 207, 225], dtype=torch.float32)
    pd = m(inp, offset)
    assert (gt - pd.cpu().flatten()).abs().sum().item() < 1e-8


CASES = [
    (2, 8, 9, 11, 6, 3, 1, 1, 1, 1, 4, True, True),
    (1, 8, 9, 11, 6, 3, 2, 2, 2, 2, 4, True, True),
    (1, 4, 7, 5, 4, 1, 1, 0, 1, 1, 1, True, False),
    (2, 6, 8, 8, 3, 3, 1, 1, 1, 3, 2, False, False),
    (1, 64, 33, 47, 64, 3, 1, 1, 1, 1, 16, True, True),
    (1, 16, 10, 13, 16, 3, 1, 1, 1, 1, 16, True, True),
    (1, 8, 6, 6, 144, 3, 1, 1, 1, 1, 2, True, True),      # Co > one CTA's 64 output channels
    (1, 2, 5, 5, 2, 5, 1, 2, 1, 1, 1, True, True),        # 5x5 taps
]


@pytest.mark.parametrize("case", CASES)
def test_generic_dcn_vs_oracle(cuda_dev, case):
    import cdfo_b200
    x, offset, mask, wt, b = _rand_case(case, seed=CASES.index(case))
    _, _, _, _, _, k, s, p, d, groups, dg, _, _ = case
    ref = O.dcn_forward(x.numpy(), offset.numpy(), None if mask is None else mask.numpy(), wt.numpy(),
                        None if b is None else b.numpy(), s, p, d, groups, dg)
    dev = lambda t: None if t is None else t.to(cuda_dev)
    if mask is None:
        y = cdfo_b200.deform_conv(dev(x), dev(offset), dev(wt), s, p, d, groups, dg)
    else:
        y = cdfo_b200.modulated_deform_conv(dev(x), dev(offset), dev(mask), dev(wt), dev(b), s, p, d, groups, dg)
    y = y.cpu().numpy()
    assert y.shape == ref.shape
    # fp32 in, fp32 accumulate: only summation order / FMA contraction differ
    assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_torchvision_compatible_entry(cuda_dev):
    import torchvision
    import cdfo_b200
    case = (2, 8, 9, 11, 6, 3, 1, 1, 1, 2, 4, True, True)
    x, offset, mask, wt, b = _rand_case(case, seed=11)
    ref = torchvision.ops.deform_conv2d(x, offset, wt, b, 1, 1, 1, mask)
    y = cdfo_b200.deform_conv2d(x.to(cuda_dev), offset.to(cuda_dev), wt.to(cuda_dev), b.to(cuda_dev), 1, 1, 1,
                                mask.to(cuda_dev))
    assert (y.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-2), (torch.bfloat16, 1e-1)])
def test_low_precision_io(cuda_dev, dtype, tol):
    """fp16 is a reference dtype (AT_DISPATCH_FLOATING_TYPES_AND_HALF); bf16 is new. fp32 accumulate inside."""
    import cdfo_b200
    case = (1, 16, 12, 12, 8, 3, 1, 1, 1, 1, 4, True, True)
    x, offset, mask, wt, b = _rand_case(case, seed=5)
    q = lambda t: t.to(dtype).float()
    ref = O.dcn_forward(q(x).numpy(), q(offset).numpy(), q(mask).numpy(), q(wt).numpy(), q(b).numpy(), 1, 1, 1, 1, 4)
    d = lambda t: t.to(cuda_dev, dtype)
    y = cdfo_b200.modulated_deform_conv(d(x), d(offset), d(mask), d(wt), d(b), 1, 1, 1, 1, 4)
    assert y.dtype == dtype
    assert np.abs(y.float().cpu().numpy() - ref).max() <= tol * max(1.0, np.abs(ref).max())


def test_sample_index_bit_exact(cuda_dev):
    """Integer MV-to-offset indexing: floor(h_im), floor(w_im) identical to the oracle for identical fp32
    offsets, including offsets that sit exactly on integers, just below them, and far outside the frame."""
    import cdfo_b200
    L = cdfo_b200._lib
    B, dg, k, H, W = 2, 16, 3, 24, 40
    g = torch.Generator().manual_seed(3)
    off = torch.randn(B, dg * 2 * k * k, H, W, generator=g) * 12.0
    off[:, ::5] = torch.round(off[:, ::5])                                 # exact integers
    off[:, 1::7] = torch.nextafter(torch.round(off[:, 1::7]), torch.tensor(-1e9))  # one ulp below an integer
    off[0, 3] = 1e4
    off[1, 4] = -1e4
    # MV prior: multiples of 1/128 added in fp32 like arch/SIDECVSR_our.py:3347
    mv = torch.randint(-64 * 3, 64 * 3, (B, 2, H, W), generator=g).float() / 128.0
    off = off + mv.flip(1).repeat(1, dg * k * k, 1, 1)
    x = np.zeros((B, dg, H, W), np.float32)
    w = np.zeros((1, dg, k, k), np.float32)
    _, ref_idx = O.dcn_forward(x, off.numpy(), None, w, None, 1, 1, 1, 1, dg, return_index=True)
    off_d = off.to(cuda_dev).contiguous()
    idx = torch.empty((B, dg * k * k, H, W, 2), dtype=torch.int32, device=cuda_dev)
    rc = L.lib().cdfo_dcn_sample_index(L.ptr(off_d), L.ptr(idx), B, H, W, k, k, 1, 1, 1, 1, 1, 1, dg,
                                       L.stream_ptr(cuda_dev))
    L.check(rc)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)


def test_error_behaviour_matches_reference(cuda_dev):
    import cdfo_b200
    x = torch.randn(1, 8, 6, 6)
    off = torch.zeros(1, 18, 6, 6)
    msk = torch.ones(1, 9, 6, 6)
    w = torch.randn(4, 8, 3, 3)
    # CPU tensors: NotImplementedError (ops/dcn/deform_conv.py:46-47,136-137)
    with pytest.raises(NotImplementedError):
        cdfo_b200.modulated_deform_conv(x, off, msk, w, None, 1, 1, 1, 1, 1)
    with pytest.raises(NotImplementedError):
        cdfo_b200.deform_conv(x, off, w, 1, 1, 1, 1, 1)
    d = lambda t: t.to(cuda_dev)
    # non-contiguous input: RuntimeError (deform_conv_cuda.cpp:493)
    xt = d(torch.randn(1, 6, 8, 6)).transpose(1, 2)
    with pytest.raises(RuntimeError, match="contiguous"):
        cdfo_b200.modulated_deform_conv(xt, d(off), d(msk), d(w), None, 1, 1, 1, 1, 1)
    # channel mismatch: RuntimeError (deform_conv_cuda.cpp:509-511)
    with pytest.raises(RuntimeError, match="wont match"):
        cdfo_b200.modulated_deform_conv(d(x), d(off), d(msk), d(torch.randn(4, 6, 3, 3)), None, 1, 1, 1, 1, 1)
    # 3-D input to DCNv1: ValueError (deform_conv.py:27-30)
    with pytest.raises(ValueError):
        cdfo_b200.deform_conv(d(x)[0], d(off), d(w), 1, 1, 1, 1, 1)
    # backward is out of scope and says so
    with pytest.raises(NotImplementedError):
        cdfo_b200.deform_conv_cuda.modulated_deform_conv_cuda_backward()


def test_pack_module_zero_init_is_plain_conv(cuda_dev):
    """ModulatedDeformConvPack at init: offsets 0, mask sigmoid(0)=0.5 -> 0.5 * conv2d(x, W) + b."""
    import cdfo_b200
    torch.manual_seed(0)
    m = cdfo_b200.ModulatedDeformConvPack(8, 8, 3, padding=1, deformable_groups=2).to(cuda_dev)
    m.bias.data.normal_()
    x = torch.randn(2, 8, 10, 12, device=cuda_dev)
    y = m(x)
    ref = 0.5 * torch.nn.functional.conv2d(x.cpu(), m.weight.cpu(), None, 1, 1) + m.bias.cpu().view(1, -1, 1, 1)
    assert (y.cpu() - ref).abs().max().item() < 1e-4


def test_torchvision_override_routes_cuda_calls_here(cuda_dev):
    """torch.library CUDA-key override of torchvision::deform_conv2d: the reference's call sites
    (arch/SIDECVSR_our.py:3164,3260,3352,3733) reach this library with zero changes to arch.py."""
    import torchvision
    import cdfo_b200
    from cdfo_b200 import torchvision_override as tvo
    case = (2, 8, 9, 11, 6, 3, 1, 1, 1, 2, 4, True, True)
    x, offset, mask, wt, b = _rand_case(case, seed=21)
    ref = O.dcn_forward(x.numpy(), offset.numpy(), mask.numpy(), wt.numpy(), b.numpy(), 1, 1, 1, 2, 4)
    ref_v1 = O.dcn_forward(x.numpy(), offset.numpy(), None, wt.numpy(), None, 1, 1, 1, 2, 4)
    d = lambda t: t.to(cuda_dev)
    tvo.install()
    try:
        n0 = cdfo_b200._lib.launch_count
        y = torchvision.ops.deform_conv2d(d(x), d(offset), d(wt), d(b), 1, 1, 1, d(mask))
        y1 = torchvision.ops.deform_conv2d(d(x), d(offset), d(wt), None, 1, 1, 1, None)
        assert cdfo_b200._lib.launch_count >= n0 + 2, "the override was not dispatched to"
        # hot shape -> the tensor-core kernel, through the same operator
        xh, oh_, mh, wh, bh = _rand_case((1, 64, 12, 16, 64, 3, 1, 1, 1, 1, 16, True, True), seed=22)
        cdfo_b200.config.tensor_core = True          # (the file's fixture pins the fp32 kernel; restored by it afterwards)
        yh = torchvision.ops.deform_conv2d(d(xh), d(oh_), d(wh), d(bh), 1, 1, 1, d(mh))
        cdfo_b200.config.tensor_core = False
        refh = O.dcn_forward(xh.numpy(), oh_.numpy(), mh.numpy(), wh.numpy(), bh.numpy(), 1, 1, 1, 1, 16)
    finally:
        tvo.uninstall()
    assert np.abs(y.cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    assert np.abs(y1.cpu().numpy() - ref_v1).max() <= 2e-5 * max(1.0, np.abs(ref_v1).max())
    assert np.abs(yh.cpu().numpy() - refh).max() <= 1e-2 * np.abs(refh).max()
    n1 = cdfo_b200._lib.launch_count
    torchvision.ops.deform_conv2d(d(x), d(offset), d(wt), d(b), 1, 1, 1, d(mask))      # uninstalled: torchvision's own kernel again
    assert cdfo_b200._lib.launch_count == n1
