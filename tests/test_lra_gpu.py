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
