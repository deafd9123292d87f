"""Feature extraction on c8 bf16 (cdfo_b200/features.py + csrc/features_c8.cu) against the oracle restatement of
PAItransformerSA_2 (oracle/torch_ref.feature_extraction, arch/SIDECVSR_our.py:1441-1475, pinned by the model goldens) and, kernel by
kernel, against plain fp32 torch on the same bf16-rounded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_util as G
from oracle import torch_ref

pytestmark = pytest.mark.gpu


def _c8(t, dev):
    from cdfo_b200 import conv
    return conv.to_c8(t.to(dev))


def _nchw(t8):
    B, C8, H, W, _ = t8.shape
    return t8.permute(0, 1, 4, 2, 3).reshape(B, C8 * 8, H, W).float().cpu()


def _q(t):
    return t.to(torch.bfloat16).float()


def _model(dev):
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8(alignment="mv_dcn")
    m.load_state_dict(G.seeded_weights("O2"), strict=True)
    return m.to(dev).eval()


def test_layernorm_dwconv_c8(cuda_dev):
    from cdfo_b200 import _lib, features
    m = _model(cuda_dev)
    g = torch.Generator().manual_seed(1)
    x = _q(torch.randn(2, 64, 20, 28, generator=g))
    norm = m.transformer_feature_extraction.path1.norm1
    y = _nchw(features.layernorm_c8(norm, _c8(x, cuda_dev)))
    mu, var = x.mean(1, keepdim=True), x.var(1, keepdim=True, unbiased=False)
    ref = (x - mu) / torch.sqrt(var + 1e-5) * norm.body.weight.detach().cpu().view(1, -1, 1, 1) + norm.body.bias.detach().cpu().view(1, -1, 1, 1)
    assert (y - ref).abs().max().item() <= 2e-2            # bf16 output rounding of values up to ~4
    x3 = _q(torch.randn(2, 192, 20, 28, generator=g))
    w = torch.randn(192, 1, 3, 3, generator=g) * 0.3
    x8 = _c8(x3, cuda_dev)
    out = torch.empty_like(x8)
    _lib.call("cdfo_dwconv3x3_c8_fwd", _lib.ptr(x8), _lib.ptr(w.reshape(192, 9).contiguous().to(cuda_dev)), _lib.ptr(out), 2, 192, 20, 28,
              _lib.stream_ptr(cuda_dev))
    ref = F.conv2d(x3, w, padding=1, groups=192)
    assert (_nchw(out) - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())


def test_self_attention_c8_vs_torch(cuda_dev):
    """qkv 1x1 -> depthwise -> Gram -> fold -> apply (+ both residual outputs) against arch:1545-1576 in fp32 torch."""
    from cdfo_b200 import features
    m = _model(cuda_dev)
    attn = m.transformer_feature_extraction.path1.attn
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    n = _q(torch.randn(2, 64, 24, 40, generator=g))
    x1 = _q(torch.randn(2, 64, 24, 40, generator=g))
    x2 = _q(torch.randn(2, 64, 24, 40, generator=g))
    o1, o2 = features.self_mdta_c8(attn, _c8(n, cuda_dev), _c8(x1, cuda_dev), _c8(x2, cuda_dev))
    ref = x1 + torch_ref._self_mdta(sd, "transformer_feature_extraction.path1.attn", n)
    e1 = (_nchw(o1) - ref).abs().max().item()
    e2 = (_nchw(o2) - (ref + x2)).abs().max().item()
    print("self attention c8: max err %.3g / %.3g (max|ref| %.3g)" % (e1, e2, ref.abs().max().item()))
    assert e1 <= 3e-2 and e2 <= 5e-2
    o1b, _ = features.self_mdta_c8(attn, _c8(n, cuda_dev), _c8(x1, cuda_dev), _c8(x2, cuda_dev))
    assert torch.equal(o1, o1b)                              # fixed reduction order


@pytest.mark.parametrize("H,W", [(24, 40), (16, 8), (40, 56)])
def test_side_branch_c8_vs_oracle(cuda_dev, H, W):
    from cdfo_b200 import features
    m = _model(cuda_dev)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(H)
    s = _q(torch.randn(2, 64, H, W, generator=g) * 0.5)
    r = _q(torch.randn(2, 64, H, W, generator=g) * 0.5)
    side = m.transformer_feature_extraction.path1.side_to_feaoneUDSA
    out = _nchw(features.side_branch_c8(side, _c8(s, cuda_dev), _c8(r, cuda_dev)))
    ref = torch_ref._side_branch(sd, "transformer_feature_extraction.path1.side_to_feaoneUDSA", s) + r
    err = (out - ref).abs().max().item()
    print("side branch c8 %dx%d: max err %.3g (max|ref| %.3g)" % (H, W, err, ref.abs().max().item()))
    assert err <= 2e-2 * max(1.0, ref.abs().max().item())


def test_feature_extraction_c8_vs_oracle_and_rerun(cuda_dev):
    from cdfo_b200 import features, synthetic
    m = _model(cuda_dev)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    clip = synthetic.make_clip(3, 40, 56, 2)
    x, pms = clip["x"][:, 3], clip["pms"][:, 3]                     # [2, 1, H, W]
    with torch.no_grad():
        l1 = F.leaky_relu(F.conv2d(x, sd["conv_first.weight"], sd["conv_first.bias"], padding=1), 0.1)
        s0 = F.conv2d(pms, sd["conv_second.weight"], sd["conv_second.bias"], padding=1)
        ref = torch_ref.feature_extraction(sd, l1, s0)
    out8 = features.feature_extraction_c8(m, x.to(cuda_dev), pms.to(cuda_dev))
    out = _nchw(out8)
    err, scale = (out - ref).abs().max().item(), ref.abs().max().item()
    rel = ((out - ref).norm() / ref.norm()).item()
    print("feature extraction c8 40x56: max err %.3g (max|ref| %.3g), relative L2 %.3g" % (err, scale, rel))
    assert err <= 3e-2 * max(1.0, scale) and rel <= 1e-2
    assert torch.equal(features.feature_extraction_c8(m, x.to(cuda_dev), pms.to(cuda_dev)), out8)    # bit-identical rerun


def test_model_uses_c8_features_without_library_kernels(cuda_dev):
    """The benchmarked configuration (lowp = bf16) must not launch cuDNN / cuBLAS kernels in the feature extraction: compare the c8 path
    against round 1's cuDNN path and check that both stay within the tolerance of each other."""
    import cdfo_b200
    from cdfo_b200 import synthetic
    m = _model(cuda_dev)
    m.lowp = torch.bfloat16
    clip = synthetic.make_clip(5, 40, 56, 1)
    x, pms = clip["x"][:, 3].to(cuda_dev), clip["pms"][:, 3].to(cuda_dev)
    assert cdfo_b200.config.features_c8
    a = m._features(x, pms)
    cdfo_b200.config.features_c8 = False
    try:
        b = m._features(x, pms)
    finally:
        cdfo_b200.config.features_c8 = True
    rel = ((a - b).norm() / b.norm()).item()
    print("c8 features vs cuDNN bf16 features: relative L2 %.3g" % rel)
    assert a.dtype == torch.float32 and a.shape == b.shape and rel <= 2e-2
