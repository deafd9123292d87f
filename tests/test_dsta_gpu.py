"""DSTA interface mirror (ops/attentionlayer.py:86-156) against the oracle restatement; state_dict names as the reference's."""
import pytest
import torch

import numpy as np

import golden_util as G
from oracle import torch_ref

pytestmark = pytest.mark.gpu

REF_KEYS = {  # parameter names a reference DSTA(64) state_dict holds (ops/attentionlayer.py:90-115)
    "conv1", "conv_f", "conv_max", "conv2", "conv3", "conv3_", "conv4", "dcn", "mask", "down_conv2.0", "mask2", "conv_du.0", "conv_du.2"}


def test_dsta_matches_oracle_and_keeps_names(cuda_dev):
    import cdfo_b200
    torch.manual_seed(3)
    m = cdfo_b200.DSTA(64)
    with torch.no_grad():
        m.dcn.bias.normal_(0, 0.1)
        m.mask.weight.mul_(3.0)          # offsets of a few pixels
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert {k.rsplit(".", 1)[0] for k in sd} == REF_KEYS
    x = torch.randn(2, 64, 96, 128)
    with torch.no_grad():
        ref = torch_ref.dsta(sd, x)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the dense convolutions around the DCN are cuDNN calls: compare in fp32
    try:
        got = m.to(cuda_dev)(x.to(cuda_dev)).cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    err = (got - ref).abs().max().item()
    print("DSTA max err %.3g (max|ref| %.3g)" % (err, ref.abs().max().item()))
    assert err <= 2e-4
    with pytest.raises(NotImplementedError):
        cdfo_b200.DSTA(64)(x)


def test_dsta_vs_reference_golden(cuda_dev):
    """A11 against the REAL reference: tests/golden/dsta_golden.npz is the output of ops/attentionlayer.py's own forward."""
    import cdfo_b200
    g = G.load("dsta_golden.npz")
    m = cdfo_b200.DSTA(64)
    sd, x = G.dsta_inputs(m.state_dict())
    m.load_state_dict(sd, strict=True)                     # the reference's parameter names / shapes
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        got = m.to(cuda_dev)(x.to(cuda_dev)).cpu().numpy()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    err = np.abs(got - g["out"]).max()
    print("DSTA vs reference golden: max err %.3g (max|ref| %.3g)" % (err, np.abs(g["out"]).max()))
    assert err <= 2e-4
