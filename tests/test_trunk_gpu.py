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
