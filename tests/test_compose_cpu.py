"""Host-side algebra of the execution plan, checked on the CPU against plain torch (no kernel involved): the weight compositions that let
hotpath.py drop launches must be exact up to fp32 rounding, borders included (Block_.forward arch/SIDECVSR_our.py:378-406, Interpolate
:324-333, LLongRangAttention arch:2183 after conv_expand_rms :4447)."""
import torch
import torch.nn.functional as F

from cdfo_b200 import hotpath


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


def test_1x1_after_3x3_composes_into_the_3x3():
    w3, b3, w1, b1 = _rand(24, 16, 3, 3, seed=1), _rand(24, seed=2), _rand(8, 24, 1, 1, seed=3), _rand(8, seed=4)
    x = _rand(2, 16, 9, 11, seed=5)
    ref = F.conv2d(F.conv2d(x, w3, b3, padding=1), w1, b1)
    w, b = hotpath._compose_1x1_after_3x3(w1, b1, w3, b3)
    got = F.conv2d(x, w.double(), b.double(), padding=1)
    assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_3x3_after_1x1_needs_the_border_class_bias():
    """conv3x3(zero padding) o conv1x1: composed weights + a bias per border class (the 1x1 bias passes only through the taps inside the
    frame) reproduce the chain on every pixel, corners included; the interior bias alone is wrong on the border ring."""
    w3, b3, w1, b1 = _rand(12, 16, 3, 3, seed=6), _rand(12, seed=7), _rand(16, 8, 1, 1, seed=8), _rand(16, seed=9, scale=3.0)
    x = _rand(2, 8, 7, 6, seed=10)
    ref = F.conv2d(F.conv2d(x, w1, b1), w3, b3, padding=1)
    w, b, be = hotpath._compose_3x3_after_1x1(w3, b3, w1, b1)
    assert tuple(be.shape) == (9, 12) and torch.equal(b, be[4])
    raw = F.conv2d(x, w.double(), None, padding=1)
    H, W = x.shape[2:]
    cls = torch.tensor([[(0 if h == 0 else (2 if h == H - 1 else 1)) * 3 + (0 if c == 0 else (2 if c == W - 1 else 1)) for c in range(W)]
                        for h in range(H)])
    got = raw + be.double()[cls].permute(2, 0, 1).unsqueeze(0)
    assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    wrong = raw + b.double().view(1, -1, 1, 1)
    assert (wrong - ref)[:, :, 1:-1, 1:-1].abs().max().item() <= 1e-5 * ref.abs().max().item()
    assert (wrong - ref)[:, :, 0, :].abs().max().item() > 0.1


def test_1x1_convolution_commutes_with_bilinear_resampling():
    """down = Interpolate(0.5) o conv1x1 and up = Interpolate(2) o conv1x1 (arch:388-399): the bilinear weights sum to 1 (edge clamping
    included), so the 1x1 may act after the resampling -- which is where hotpath.cross_scale_block composes it into body.0."""
    w1, b1 = _rand(8, 8, 1, 1, seed=11), _rand(8, seed=12)
    x = _rand(1, 8, 10, 12, seed=13)
    for scale in (0.5, 2.0):
        a = F.interpolate(F.conv2d(x, w1, b1), scale_factor=scale, mode="bilinear", align_corners=False)
        b = F.conv2d(F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=False), w1, b1)
        assert (a - b).abs().max().item() <= 1e-12 * a.abs().max().item()


def test_conv_du_re0_composes_into_the_prior_convolution():
    """v = ReLU(conv_du_re.0(conv_expand_rms(rms))): one direct 1 -> 64 convolution of the one-channel map with w' = W1 w, b' = W1 b + b1
    (the 1x1 comes AFTER the zero-padded 3x3: exact on every pixel)."""
    wp, bp, w1, b1 = _rand(64, 1, 3, 3, seed=14), _rand(64, seed=15), _rand(64, 64, 1, 1, seed=16, scale=0.2), _rand(64, seed=17)
    rms = _rand(2, 1, 9, 13, seed=18)
    ref = F.relu(F.conv2d(F.conv2d(rms, wp, bp, padding=1), w1, b1))
    m = w1.reshape(64, 64)
    got = F.relu(F.conv2d(rms, (m @ wp.reshape(64, 9)).reshape(64, 1, 3, 3), m @ bp + b1, padding=1))
    assert (got - ref).abs().max().item() <= 1e-12 * ref.abs().max().item()


def test_folded_half_convolution_weights():
    """bilinear x0.5 (2x2 mean) of a 3x3 convolution = one 4x4 / stride-2 convolution with W4[u][v] = 1/4 sum of the 3x3 taps that reach
    (u, v) from the four pixels of the block (csrc/conv4x4s2_pair_sm100.cu pack_weight_4x4_kernel restated)."""
    w3, b3 = _rand(6, 5, 3, 3, seed=19), _rand(6, seed=20)
    x = _rand(2, 5, 12, 16, seed=21)
    ref = F.interpolate(F.conv2d(x, w3, b3, padding=1), scale_factor=0.5, mode="bilinear", align_corners=False)
    w4 = torch.zeros(6, 5, 4, 4, dtype=torch.float64)
    for dy in range(2):
        for dx in range(2):
            w4[:, :, dy:dy + 3, dx:dx + 3] += 0.25 * w3
    got = F.conv2d(x, w4, b3, stride=2, padding=1)
    assert (got - ref).abs().max().item() <= 1e-12 * ref.abs().max().item()
