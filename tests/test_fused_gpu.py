"""The fused head + DCN kernel (csrc/mv_dcn_fused_sm100.cu: conv_offset[-1] on both hidden maps, tanh / sum / sigmoid, MV prior and the
deformable convolution in ONE launch, arch/SIDECVSR_our.py:3339-3352) against the two-kernel path it replaces (dual head launch ->
fp16 fields in HBM -> texture-gather DCN), which tests/test_model_gpu.py pins to the real reference's goldens.  Both evaluate the
head with the same explicit fp32 arithmetic (csrc/mv_head_math.cuh) and the same MMA accumulation order, so the debug tap of the
fields and the outputs must agree BIT FOR BIT; the oracle comparison of the module is in test_model_gpu.py (the model routes
through the fused kernel by default)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import c_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(B, H, W, xB, seed, dev):
    g = torch.Generator().manual_seed(seed)
    z = (torch.randn(2 * B, 64, H, W, generator=g) * 0.5)
    c2 = torch.nn.Conv2d(64, 432, 3, 1, 1)
    with torch.no_grad():
        c2.weight.copy_(torch.randn(432, 64, 3, 3, generator=g) * 0.03)      # offsets of a few pixels, masks spread over (0, 1)
        c2.bias.copy_(torch.randn(432, generator=g) * 0.2)
    x = torch.randn(xB, 64, H, W, generator=g)
    mv = torch.randint(-192, 192, (B, 2, H, W), generator=g).float() / 128.0
    wd = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    bd = torch.randn(64, generator=g)
    return z.to(dev), c2.to(dev), x.to(dev), mv.to(dev), wd.to(dev), bd.to(dev)


def _two_kernel(z8, c2, xq, mv, wd16, bd, B, H, W, out_c8=False):
    from cdfo_b200 import _lib, dcn_sm100 as S, hotpath
    wpk, bias = hotpath._head_weights(c2, 16)
    fields = torch.empty(S.fields_shape(B, 16, H, W), dtype=torch.float16, device=z8.device)
    _lib.call("cdfo_mv_offset_head_dual_sm100_fwd", _lib.ptr(z8), _lib.ptr(wpk), _lib.ptr(bias), _lib.ptr(fields), B, 64, 16, H, W,
              ctypes.c_float(10.0), _lib.stream_ptr(z8.device))
    return fields, S.dcn_tex(xq, fields, wd16, bd, mv=mv, out_c8=out_c8)


@pytest.mark.parametrize("B,H,W,xB", [(1, 16, 8, 1), (1, 32, 24, 1), (2, 24, 40, 1), (3, 40, 56, 3), (6, 20, 44, 1), (1, 64, 64, 1)])
def test_fused_equals_two_kernel_path_bitwise(cuda_dev, B, H, W, xB):
    from cdfo_b200 import conv, dcn_sm100 as S, hotpath
    z, c2, x, mv, wd, bd = _inputs(B, H, W, xB, 100 + H * W + B, cuda_dev)
    z8, xq, wd16 = conv.to_c8(z), S.pack_q4t(x), S.pack_weight_f16(wd)
    fields, y_ref = _two_kernel(z8, c2, xq, mv, wd16, bd, B, H, W)
    hw, hb = hotpath._head_weights_fused(c2, 16)
    tap = torch.zeros_like(fields)
    y = S.mv_head_dcn_fused(z8, hw, hb, 10.0, xq, mv, wd16, bd, fields_out=tap)
    torch.cuda.synchronize()
    nf = int((tap.view(torch.int16) != fields.view(torch.int16)).sum())
    ny = int((y != y_ref).sum())
    print("fused vs two-kernel B%d %dx%d: %d / %d field words differ, %d / %d outputs differ (max |dy| %.3g)"
          % (B, H, W, nf, fields.numel(), ny, y.numel(), float((y - y_ref).abs().max())))
    assert nf == 0 and ny == 0
    # c8 output = the same values rounded to bf16
    y8 = S.mv_head_dcn_fused(z8, hw, hb, 10.0, xq, mv, wd16, bd, out_c8=True)
    assert torch.equal(y8.permute(0, 1, 4, 2, 3).reshape(B, 64, H, W).float(), y_ref.to(torch.bfloat16).float())


def test_fused_stacked_output_and_rerun(cuda_dev):
    """Group-major batch written straight into tsa_fusion's stacked input (what the model uses) == the texture kernel's stacked
    variant; a second launch gives identical bits (fixed reduction order, no atomics)."""
    from cdfo_b200 import conv, dcn_sm100 as S, hotpath
    n_seq, n_grp, H, W = 2, 3, 24, 40
    B = n_seq * n_grp
    z, c2, x, mv, wd, bd = _inputs(B, H, W, n_seq, 7, cuda_dev)
    z8, xq, wd16 = conv.to_c8(z), S.pack_q4t(x), S.pack_weight_f16(wd)
    wpk, bias = hotpath._head_weights(c2, 16)
    fields, _ = _two_kernel(z8, c2, xq, mv, wd16, bd, B, H, W)
    chunks = [0, 16, 40]
    ref = torch.zeros((n_seq, 56, H, W, 8), dtype=torch.bfloat16, device=cuda_dev)
    S.dcn_tex_stacked(xq, fields, wd16, bd, mv, ref, chunks)
    hw, hb = hotpath._head_weights_fused(c2, 16)
    out = torch.zeros_like(ref)
    S.mv_head_dcn_fused(z8, hw, hb, 10.0, xq, mv, wd16, bd, stack=out, group_chunk=chunks)
    assert torch.equal(out, ref)
    out2 = torch.zeros_like(ref)
    S.mv_head_dcn_fused(z8, hw, hb, 10.0, xq, mv, wd16, bd, stack=out2, group_chunk=chunks)
    assert torch.equal(out2, out)


def test_fused_vs_oracle(cuda_dev):
    """Independent of the two-kernel path: head in fp32 torch on the bf16-rounded operands -> offsets / mask -> C oracle DCN."""
    from cdfo_b200 import conv, dcn_sm100 as S, hotpath
    B, H, W = 2, 24, 40
    z, c2, x, mv, wd, bd = _inputs(B, H, W, 1, 3, cuda_dev)
    z8, xq, wd16 = conv.to_c8(z), S.pack_q4t(x), S.pack_weight_f16(wd)
    hw, hb = hotpath._head_weights_fused(c2, 16)
    y = S.mv_head_dcn_fused(z8, hw, hb, 10.0, xq, mv, wd16, bd).cpu().numpy()
    zq = z.to(torch.bfloat16).float().cpu()
    wq = c2.weight.detach().to(torch.bfloat16).float().cpu()
    o = [torch.nn.functional.conv2d(zq[k * B:(k + 1) * B], wq, c2.bias.detach().cpu(), padding=1) for k in range(2)]
    off = 10.0 * torch.tanh(o[0][:, :288]) + 10.0 * torch.tanh(o[1][:, :288]) + mv.cpu().flip(1).repeat(1, 144, 1, 1)   # arch:3345-3347
    msk = torch.sigmoid(o[0][:, 288:] + o[1][:, 288:])                                                                   # arch:3350
    xr = x.half().float().cpu().repeat(B, 1, 1, 1)
    ref = O.dcn_forward(xr.numpy(), off.numpy(), msk.numpy(), wd.half().float().cpu().numpy(), bd.cpu().numpy(), 1, 1, 1, 1, 16)
    err, scale = float(np.abs(y - ref).max()), float(np.abs(ref).max())
    print("fused head + DCN vs torch head + C oracle DCN: max err %.3g (max|ref| %.3g)" % (err, scale))
    assert err <= 2e-2 * scale
