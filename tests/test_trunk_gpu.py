"""Resample kernels and the c8 reconstruction trunk against torch / the oracle."""
import pytest
import torch
import torch.nn.functional as F

import golden_util as G
from oracle import torch_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W", [(8, 8), (24, 40), (34, 50)])
def test_resample_c8_vs_interpolate(cuda_dev, H, W):
    from cdfo_b200 import conv
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(2, 64, H, W, generator=g).to(torch.bfloat16).float().to(cuda_dev)
    x8 = conv.to_c8(x)
    up = conv.from_c8(conv.resample(x8, 1))
    ref_up = F.interpolate(x, scale_factor=2.0, mode="bilinear", align_corners=False)
    assert (up - ref_up).abs().max().item() <= 2 ** -8 * ref_up.abs().max().item()      # bf16 output rounding only
    dn = conv.from_c8(conv.resample(x8, 0))
    ref_dn = F.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=False)
    assert (dn - ref_dn).abs().max().item() <= 2 ** -8 * ref_dn.abs().max().item()
    big = torch.randn(2, 64, 2 * H, 2 * W, generator=g).to(torch.bfloat16).float().to(cuda_dev)
    small = torch.randn(2, 64, H // 2, W // 2, generator=g).to(torch.bfloat16).float().to(cuda_dev)
    got = conv.from_c8(conv.resample(conv.to_c8(big), 2, b=conv.to_c8(small), base=x8))
    ref = x + F.interpolate(big, scale_factor=0.5, mode="bilinear", align_corners=False) + \
        F.interpolate(small, scale_factor=2.0, mode="bilinear", align_corners=False)
    assert (got - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()


def test_trunk_vs_oracle(cuda_dev):
    """SCNet_ (21 cross-scale blocks) on c8 bf16 with composed 1x1 convolutions vs the fp32 oracle (arch:378-480)."""
    from cdfo_b200 import conv, hotpath
    from cdfo_b200.model import CVSR_V8
    sd = G.seeded_weights("O1")
    m = CVSR_V8()
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda_dev).eval()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 64, 24, 40, generator=g)
    with torch.no_grad():
        ref = torch_ref.trunk(sd, x)
    got = conv.from_c8(hotpath.recon_trunk(m.recon_trunk, conv.to_c8(x.to(cuda_dev)))).cpu()
    err = (got - ref).abs().max().item()
    print("trunk max err %.3g (max|ref| %.3g)" % (err, ref.abs().max().item()))
    assert err <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_and_dwconv_vs_torch(cuda_dev, dtype):
    """Channel LayerNorm (arch:1169-1198) and the depthwise 3x3 of the MDTA attention (arch:1545-1576)."""
    from cdfo_b200 import hotpath
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 64, 19, 37, generator=g).to(dtype)
    gamma, beta = torch.randn(64, generator=g), torch.randn(64, generator=g)
    xf = x.float()
    mu, var = xf.mean(1, keepdim=True), xf.var(1, keepdim=True, unbiased=False)
    ref = (xf - mu) * torch.rsqrt(var + 1e-5) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    got = hotpath.layernorm_c(x.to(cuda_dev), gamma.to(cuda_dev), beta.to(cuda_dev)).float().cpu()
    tol = 1e-5 if dtype == torch.float32 else 2 ** -7
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()
    x3 = torch.randn(2, 192, 19, 37, generator=g).to(dtype)
    w = torch.randn(192, 1, 3, 3, generator=g)
    ref = F.conv2d(x3.float(), w, None, 1, 1, 1, 192)
    got = hotpath.dwconv3x3(x3.to(cuda_dev), w.to(cuda_dev)).float().cpu()
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()
    x4 = torch.randn(1, 24, 11, 40, generator=g).to(dtype)            # W % 8 == 0: vectorised bf16 path
    w4 = torch.randn(24, 1, 3, 3, generator=g)
    ref = F.conv2d(x4.float(), w4, None, 1, 1, 1, 24)
    got = hotpath.dwconv3x3(x4.to(cuda_dev), w4.to(cuda_dev)).float().cpu()
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item()
