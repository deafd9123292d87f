"""MDTA kernels (csrc/mdta.cu) against the torch oracle: both alignment variants, a ragged pixel count (H*W not a
multiple of the 128-pixel tile), a query shared by several samples, flows that leave the frame."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_util as G
from oracle import torch_ref

pytestmark = pytest.mark.gpu


def _inputs(B, xB, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(xB, 64, H, W, generator=g)
    extra = torch.randn(B, 64, H, W, generator=g)
    pred = torch.randn(B, 64, H, W, generator=g) * 0.5
    flow = torch.randint(-64 * 3, 64 * 3, (B, 2, (H + 7) // 8, (W + 7) // 8), generator=g).float() / 128.0
    flow = flow.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :H, :W].contiguous()
    flow[0, :, 0, :4] = 40.0          # far outside the frame
    return x, extra, pred, flow


@pytest.mark.parametrize("B,xB,H,W", [(2, 1, 16, 24), (3, 3, 19, 23), (6, 2, 32, 40)])
def test_mdta_mode0_vs_oracle(cuda_dev, B, xB, H, W):
    """MVDualAttAlignment front end (arch:3303-3337): o1, o2 = project_out(attn @ (v * gate)); bf16 c8 output."""
    from cdfo_b200 import conv, hotpath
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8(alignment="mv_dcn")
    sd = G.seeded_weights("O2")
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda_dev).eval()
    x, extra, pred, flow = _inputs(B, xB, H, W, seed=B * 100 + H)
    pre = "MV_deform_align."
    xr = x.repeat(B // xB, 1, 1, 1)
    with torch.no_grad():
        warped = torch_ref.flow_warp(extra, flow.permute(0, 2, 3, 1))
        fused = F.conv2d(torch.cat([warped, pred], 1), sd[pre + "fusion_out.weight"])
        t = sd[pre + "temperature"]
        o1 = F.conv2d(torch_ref._mdta(xr, fused, warped * torch_ref._channel_gate(sd, pre + "conv_du", warped), t, 8), sd[pre + "project_out.weight"])
        o2 = F.conv2d(torch_ref._mdta(xr, fused, pred * torch_ref._channel_gate(sd, pre + "conv_du", pred), t, 8), sd[pre + "project_out.weight"])
    d = lambda v: v.to(cuda_dev)
    z = hotpath.dual_mdta(m.MV_deform_align, d(x), d(extra), d(pred), d(flow), mode=0)
    got = conv.from_c8(z).cpu()
    ref = torch.cat([o1, o2], 0)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print("mdta mode0 B%d xB%d %dx%d: max err %.3g (max|ref| %.3g)" % (B, xB, H, W, err, scale))
    assert err <= 6e-3 * scale + 1e-5          # bf16 output rounding (2^-9) dominates


@pytest.mark.parametrize("B,xB,H,W", [(2, 1, 16, 24), (3, 3, 19, 23), (6, 2, 32, 40), (12, 2, 72, 120)])
def test_mdta_mode0_c8_inputs_vs_oracle(cuda_dev, B, xB, H, W):
    """The same front end fed with c8 bf16 tensors (cdfo_mdta_c8_fwd: what conv_expand_fea_r, the prior convolution and the stack pack
    write), against the oracle evaluated on the SAME bf16-rounded inputs, and against the fp32-input kernels."""
    from cdfo_b200 import conv, hotpath
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8(alignment="mv_dcn")
    sd = G.seeded_weights("O2")
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda_dev).eval()
    x, extra, pred, flow = _inputs(B, xB, H, W, seed=B * 100 + H + 1)
    x, extra, pred = (t.to(torch.bfloat16).float() for t in (x, extra, pred))
    pre = "MV_deform_align."
    xr = x.repeat(B // xB, 1, 1, 1)
    with torch.no_grad():
        warped = torch_ref.flow_warp(extra, flow.permute(0, 2, 3, 1))
        fused = F.conv2d(torch.cat([warped, pred], 1), sd[pre + "fusion_out.weight"])
        t = sd[pre + "temperature"]
        o1 = F.conv2d(torch_ref._mdta(xr, fused, warped * torch_ref._channel_gate(sd, pre + "conv_du", warped), t, 8), sd[pre + "project_out.weight"])
        o2 = F.conv2d(torch_ref._mdta(xr, fused, pred * torch_ref._channel_gate(sd, pre + "conv_du", pred), t, 8), sd[pre + "project_out.weight"])
    d = lambda v: v.to(cuda_dev)
    z = hotpath.dual_mdta(m.MV_deform_align, conv.to_c8(d(x)), conv.to_c8(d(extra)), conv.to_c8(d(pred)), d(flow), mode=0)
    z32 = hotpath.dual_mdta(m.MV_deform_align, d(x), d(extra), d(pred), d(flow), mode=0)
    got, got32 = conv.from_c8(z).cpu(), conv.from_c8(z32).cpu()
    ref = torch.cat([o1, o2], 0)
    err, dd = (got - ref).abs().max().item(), (got - got32).abs().max().item()
    scale = ref.abs().max().item()
    print("mdta mode0 c8 inputs B%d xB%d %dx%d: max err %.3g, vs fp32-input kernels %.3g (max|ref| %.3g)" % (B, xB, H, W, err, dd, scale))
    assert err <= 8e-3 * scale + 1e-5          # bf16 rounding of the warped features (2^-9) on top of the output rounding
    assert dd <= 8e-3 * scale + 1e-5
    assert torch.equal(hotpath.dual_mdta(m.MV_deform_align, conv.to_c8(d(x)), conv.to_c8(d(extra)), conv.to_c8(d(pred)), d(flow), mode=0), z)


@pytest.mark.parametrize("B,xB,H,W", [(2, 1, 16, 24), (3, 3, 19, 23)])
def test_dual_att_alignment_vs_oracle(cuda_dev, B, xB, H, W):
    """Whole DualAttAlignment (arch:3455-3496): MDTA kernels (fp32) + CALayer gate + bf16 tcgen05 residual blocks."""
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8(alignment="dual_att")
    sd = G.seeded_weights("O1")
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda_dev).eval()
    x, extra, pred, flow = _inputs(B, xB, H, W, seed=7 + H)
    with torch.no_grad():
        ref = torch_ref.dual_att_alignment(sd, "MV_deform_align.", x.repeat(B // xB, 1, 1, 1), extra, pred, flow)
    d = lambda v: v.to(cuda_dev)
    got = m.MV_deform_align(d(x), d(extra), d(pred), d(flow)).cpu()
    err = (got - ref).abs().max().item()
    print("DualAttAlignment B%d xB%d %dx%d: max err %.3g (max|ref| %.3g)" % (B, xB, H, W, err, ref.abs().max().item()))
    assert err <= 2e-2 * ref.abs().max().item()


def test_mdta_is_deterministic(cuda_dev):
    from cdfo_b200 import hotpath
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8(alignment="mv_dcn")
    m.load_state_dict(G.seeded_weights("O2"), strict=True)
    m = m.to(cuda_dev).eval()
    x, extra, pred, flow = [t.to(cuda_dev) for t in _inputs(4, 2, 40, 56, seed=3)]
    a = hotpath.dual_mdta(m.MV_deform_align, x, extra, pred, flow, mode=0)
    b = hotpath.dual_mdta(m.MV_deform_align, x, extra, pred, flow, mode=0)
    assert torch.equal(a, b)
