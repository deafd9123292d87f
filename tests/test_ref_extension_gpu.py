"""Same-box parity against the REFERENCE'S OWN compiled CUDA kernel (ops/dcn/src/deform_conv_cuda_kernel.cu:570-632 +
deform_conv_cuda.cpp:486-564, built unmodified but for the 6 `.type()` -> `.scalar_type()` replacements by baseline/build_ref_dcn.py into
baseline/_ref/): both modules are called through the identical pybind-style signature the reference's deform_conv.py uses."""
import glob
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

_SO = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "deform_conv_cuda*.so")))


def _ref_ext():
    spec = importlib.util.spec_from_file_location("deform_conv_cuda", _SO[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not _SO, reason="baseline/_ref/deform_conv_cuda*.so not built (python baseline/build_ref_dcn.py in the build container)")
@pytest.mark.parametrize("case", [
    # B, C, H, W, Co, k, stride, pad, dil, groups, dg
    (2, 64, 24, 40, 64, 3, 1, 1, 1, 1, 16),      # the model's hot shape (A5/A6)
    (1, 16, 43, 78, 16, 3, 1, 1, 1, 1, 16),      # DSTA's internal DCN at c3 (A11)
    (2, 8, 9, 11, 6, 3, 2, 1, 1, 2, 4),
    (1, 4, 12, 12, 4, 3, 1, 2, 2, 1, 1),
])
def test_modulated_dcn_equals_reference_cuda_kernel(cuda_dev, case):
    import cdfo_b200
    import cdfo_b200.deform_conv_cuda as ours
    ref = _ref_ext()
    B, C, H, W, Co, k, s, p, d, groups, dg = case
    g = torch.Generator().manual_seed(sum(case))
    Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1
    Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    x = torch.randn(B, C, H, W, generator=g).to(cuda_dev)
    off = (torch.randn(B, dg * 2 * k * k, Ho, Wo, generator=g) * 3.0).to(cuda_dev)
    off[:, ::4] = torch.round(off[:, ::4])
    msk = torch.rand(B, dg * k * k, Ho, Wo, generator=g).to(cuda_dev)
    w = (torch.randn(Co, C // groups, k, k, generator=g) * 0.2).to(cuda_dev)
    b = torch.randn(Co, generator=g).to(cuda_dev)
    saved = cdfo_b200.config.tensor_core
    cdfo_b200.config.tensor_core = False
    try:
        outs = []
        for mod in (ref, ours):
            out = x.new_empty((B, Co, Ho, Wo))
            bufs = [x.new_empty(0), x.new_empty(0)]
            mod.modulated_deform_conv_cuda_forward(x, w, b, bufs[0], off, msk, out, bufs[1], k, k, s, s, p, p, d, d, groups, dg, True)   # deform_conv.py:144-148
            outs.append(out)
        torch.cuda.synchronize()
    finally:
        cdfo_b200.config.tensor_core = saved
    err = (outs[0] - outs[1]).abs().max().item()
    scale = outs[0].abs().max().item()
    print("reference CUDA kernel vs cdfo_dcn_fwd %s: max |diff| %.3g (max|ref| %.3g)" % (case, err, scale))
    assert err <= 2e-5 * max(1.0, scale)          # both fp32; the reference sums through im2col + cuBLAS SGEMM, ours in a fixed order


@pytest.mark.skipif(not _SO, reason="baseline/_ref not built")
def test_dcn_v1_equals_reference_cuda_kernel(cuda_dev):
    import cdfo_b200.deform_conv_cuda as ours
    ref = _ref_ext()
    B, C, H, W, Co, k, dg = 2, 8, 10, 14, 6, 3, 2
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, H, W, generator=g).to(cuda_dev)
    off = (torch.randn(B, dg * 18, H, W, generator=g) * 2.0).to(cuda_dev)
    w = (torch.randn(Co, C, k, k, generator=g) * 0.2).to(cuda_dev)
    outs = []
    for mod in (ref, ours):
        out = x.new_empty((B, Co, H, W))
        bufs = [x.new_empty(0), x.new_empty(0)]
        rc = mod.deform_conv_forward_cuda(x, w, off, out, bufs[0], bufs[1], k, k, 1, 1, 1, 1, 1, 1, 1, dg, 2)      # deform_conv.py:52-57
        assert rc == 1
        outs.append(out)
    torch.cuda.synchronize()
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-5 * max(1.0, outs[0].abs().max().item())
