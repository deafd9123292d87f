"""tcgen05 path: plumbing self-test, then the DCN implicit-GEMM kernel against the pinned C oracle."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as O

pytestmark = pytest.mark.gpu


def test_umma_selftest(cuda_dev):
    import cdfo_b200
    L = cdfo_b200._lib
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 64, generator=g).to(torch.bfloat16)
    Bm = torch.randn(64, 64, generator=g).to(torch.bfloat16)
    ref = A.float() @ Bm.float().t()
    Ad, Bd = A.to(cuda_dev), Bm.to(cuda_dev)
    D = torch.zeros(128, 64, device=cuda_dev)
    L.check(L.lib().cdfo_umma_selftest(L.ptr(Ad), L.ptr(Bd), L.ptr(D), 0, L.stream_ptr(cuda_dev)))
    torch.cuda.synchronize()
    err = (D.cpu() - ref).abs().max().item()
    assert err < 1e-3, "tcgen05 descriptor convention is wrong: max err %g" % err


def _case(B, H, W, dg, seed, off_scale=3.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 64, H, W, generator=g).to(torch.bfloat16).float()
    offset = torch.randn(B, dg * 18, H, W, generator=g) * off_scale
    mask = torch.rand(B, dg * 9, H, W, generator=g)
    wt = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).to(torch.bfloat16).float()
    b = torch.randn(64, generator=g)
    return x, offset, mask, wt, b


@pytest.mark.parametrize("B,H,W,dg", [(1, 16, 24, 16), (2, 33, 47, 16), (1, 64, 64, 16), (1, 20, 20, 4), (3, 8, 8, 1)])
def test_dcn_sm100_vs_oracle(cuda_dev, B, H, W, dg):
    """x and W are pre-rounded to bf16 on both sides, so the only differences are the bf16 rounding of the
    blended A operand (rel 2^-9 per term) and the accumulation order."""
    from cdfo_b200 import dcn_sm100 as S
    x, offset, mask, wt, b = _case(B, H, W, dg, seed=H * W + dg)
    ref = O.dcn_forward(x.numpy(), offset.numpy(), mask.numpy(), wt.numpy(), b.numpy(), 1, 1, 1, 1, dg)
    d = lambda t: t.to(cuda_dev)
    y = S.dcn_sm100(S.pack_q4p(d(x)), d(offset), d(mask), S.pack_weight(d(wt)), d(b))
    err = np.abs(y.cpu().numpy() - ref).max()
    scale = np.abs(ref).max()
    print("dcn_sm100 B%d %dx%d dg%d: max err %.3g (max|ref| %.3g)" % (B, H, W, dg, err, scale))
    assert err <= 4e-3 * scale


def test_dcn_sm100_mv_prior_and_c8_output(cuda_dev):
    """MV prior added in-kernel == offset + flow.flip(1).repeat(...) (arch/SIDECVSR_our.py:3347); c8 bf16 output."""
    from cdfo_b200 import dcn_sm100 as S
    B, H, W, dg = 1, 24, 40, 16
    x, offset, mask, wt, b = _case(B, H, W, dg, seed=1)
    g = torch.Generator().manual_seed(9)
    flow = torch.randint(-64 * 3, 64 * 3, (B, 2, H, W), generator=g).float() / 128.0
    full = offset + flow.flip(1).repeat(1, dg * 9, 1, 1)
    ref = O.dcn_forward(x.numpy(), full.numpy(), mask.numpy(), wt.numpy(), b.numpy(), 1, 1, 1, 1, dg)
    d = lambda t: t.to(cuda_dev)
    xc = S.pack_q4p(d(x))
    y = S.dcn_sm100(xc, d(offset), d(mask), S.pack_weight(d(wt)), d(b), mv=d(flow))
    y_full = S.dcn_sm100(xc, d(full), d(mask), S.pack_weight(d(wt)), d(b))
    assert torch.equal(y, y_full)          # same fp32 op order as the reference's offset assembly -> identical
    assert np.abs(y.cpu().numpy() - ref).max() <= 4e-3 * np.abs(ref).max()
    y8 = S.dcn_sm100(xc, d(offset), d(mask), S.pack_weight(d(wt)), d(b), mv=d(flow), out_c8=True)
    y8_nchw = y8.permute(0, 1, 4, 2, 3).reshape(B, 64, H, W).float()
    assert (y8_nchw - y.to(torch.bfloat16).float()).abs().max().item() == 0.0


def test_dcn_sm100_fp16_offsets_and_borders(cuda_dev):
    from cdfo_b200 import dcn_sm100 as S
    B, H, W, dg = 1, 17, 29, 16
    x, offset, mask, wt, b = _case(B, H, W, dg, seed=2, off_scale=12.0)   # many samples leave the frame
    offset[:, ::3] = torch.round(offset[:, ::3])
    offset[0, 5] = 1e4
    offset[0, 6] = float("nan")
    off16, m16 = offset.half(), mask.half()
    ref = O.dcn_forward(x.numpy(), off16.float().numpy(), m16.float().numpy(), wt.numpy(), b.numpy(), 1, 1, 1, 1, dg)
    d = lambda t: t.to(cuda_dev)
    y = S.dcn_sm100(S.pack_q4p(d(x)), d(off16), d(m16), S.pack_weight(d(wt)), d(b))
    assert np.isfinite(y.cpu().numpy()).all()
    assert np.abs(y.cpu().numpy() - ref).max() <= 4e-3 * np.abs(ref).max()


def test_dcn_sm100_full_size_linearity(cuda_dev):
    """BASELINE config c3 size (272x480): size-independent properties instead of the (slow) oracle --
    linearity in x, zero offsets + mask 1 == plain convolution, and agreement with the generic fp32 kernel."""
    import cdfo_b200
    from cdfo_b200 import dcn_sm100 as S
    B, H, W, dg = 1, 272, 480, 16
    g = torch.Generator().manual_seed(4)
    x1 = torch.randn(B, 64, H, W, generator=g).to(torch.bfloat16).float().to(cuda_dev)
    offset = (torch.randn(B, dg * 18, H, W, generator=g) * 2.0).to(cuda_dev)
    mask = torch.rand(B, dg * 9, H, W, generator=g).to(cuda_dev)
    wt = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).to(torch.bfloat16).float().to(cuda_dev)
    wpk = S.pack_weight(wt)
    y1 = S.dcn_sm100(S.pack_q4p(x1), offset, mask, wpk)
    y2 = S.dcn_sm100(S.pack_q4p(2.0 * x1), offset, mask, wpk)
    assert (y2 - 2.0 * y1).abs().max().item() == 0.0                 # scaling by 2 is exact in bf16/fp32
    gen = cdfo_b200.dcn._generic_modulated(x1, offset, mask, wt, None, 1, 1, 1, 1, dg)
    assert (y1 - gen).abs().max().item() <= 4e-3 * gen.abs().max().item()
    y0 = S.dcn_sm100(S.pack_q4p(x1), torch.zeros_like(offset), torch.ones_like(mask), wpk)
    conv = torch.nn.functional.conv2d(x1, wt, None, 1, 1)
    assert (y0 - conv).abs().max().item() <= 2e-3 * conv.abs().max().item()


def _fields(offset, mask, dg):
    """Reference-layout offset / mask -> packed fp16 fields [B, 9, H, W, dg, 4]."""
    from cdfo_b200 import dcn_sm100 as S
    return S.pack_fields(offset, mask, dg)


def _ref_off_mask(fields):
    """The fp16-rounded offset / mask the kernel really sees, back in the reference layout."""
    from cdfo_b200 import dcn_sm100 as S
    return S.unpack_fields(fields)


@pytest.mark.parametrize("B,H,W,dg", [(1, 16, 24, 16), (2, 33, 47, 16), (1, 64, 64, 16), (1, 20, 20, 4), (3, 8, 8, 1), (2, 6, 70, 16)])
def test_dcn_tex_vs_oracle(cuda_dev, B, H, W, dg):
    """Texture-unit gather: x and W pre-rounded to fp16, offsets/mask to fp16 on both sides; what remains is the
    texture filter's 8-bit weights (<= 2^-9 of the local texel differences per sample), the fp16 rounding of the
    A operand and the accumulation order."""
    from cdfo_b200 import dcn_sm100 as S
    x, offset, mask, wt, b = _case(B, H, W, dg, seed=H * W + dg)
    x, wt = x.half().float(), wt.half().float()
    fields = _fields(offset, mask, dg)
    o16, m16 = _ref_off_mask(fields)
    ref = O.dcn_forward(x.numpy(), o16.numpy(), m16.numpy(), wt.numpy(), b.numpy(), 1, 1, 1, 1, dg)
    d = lambda t: t.to(cuda_dev)
    y = S.dcn_tex(S.pack_q4t(d(x)), d(fields), S.pack_weight_f16(d(wt)), d(b))
    err = np.abs(y.cpu().numpy() - ref).max()
    scale = np.abs(ref).max()
    print("dcn_tex B%d %dx%d dg%d: max err %.3g (max|ref| %.3g)" % (B, H, W, dg, err, scale))
    assert err <= 6e-3 * scale


def test_dcn_tex_mv_borders_shared_x_c8(cuda_dev):
    """MV prior, samples far outside the frame / NaN offsets, x shared by the batch (x_batch < B), c8 output."""
    from cdfo_b200 import dcn_sm100 as S
    B, H, W, dg = 4, 17, 29, 16
    x, offset, mask, wt, b = _case(B, H, W, dg, seed=3, off_scale=10.0)
    x, wt = x[:2].half().float(), wt.half().float()
    offset[:, ::3] = torch.round(offset[:, ::3])
    offset[0, 5] = 1e4
    offset[1, 6] = float("nan")      # dy of (group 0, tap 3)
    offset[1, 9] = float("nan")      # dx of (group 0, tap 4)
    offset[2, 11] = -1e4
    g = torch.Generator().manual_seed(9)
    flow = torch.randint(-64 * 3, 64 * 3, (B, 2, H, W), generator=g).float() / 128.0
    fields = _fields(offset, mask, dg)
    o16, m16 = _ref_off_mask(fields)
    full = o16 + flow.flip(1).repeat(1, dg * 9, 1, 1)
    xb = x.repeat(2, 1, 1, 1)                      # output sample b reads x[b % 2]
    ref = O.dcn_forward(xb.numpy(), full.numpy(), m16.numpy(), wt.numpy(), b.numpy(), 1, 1, 1, 1, dg)
    d = lambda t: t.to(cuda_dev)
    xq = S.pack_q4t(d(x))
    y = S.dcn_tex(xq, d(fields), S.pack_weight_f16(d(wt)), d(b), mv=d(flow))
    assert np.isfinite(y.cpu().numpy()).all()
    err = np.abs(y.cpu().numpy() - ref).max()
    assert err <= 6e-3 * np.abs(ref).max(), err
    y8 = S.dcn_tex(xq, d(fields), S.pack_weight_f16(d(wt)), d(b), mv=d(flow), out_c8=True)
    y8_nchw = y8.permute(0, 1, 4, 2, 3).reshape(B, 64, H, W).float()
    assert (y8_nchw - y.to(torch.bfloat16).float()).abs().max().item() == 0.0


def test_dcn_tex_full_size_properties(cuda_dev):
    """272x480 (config c3): zero offsets + mask 1 == plain convolution; integer offsets are filtered exactly
    (weights 0/1); agreement with the exact-arithmetic tcgen05 kernel on random fields."""
    from cdfo_b200 import dcn_sm100 as S
    B, H, W, dg = 1, 272, 480, 16
    g = torch.Generator().manual_seed(4)
    x1 = torch.randn(B, 64, H, W, generator=g).half().float().to(cuda_dev)
    wt = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).half().float().to(cuda_dev)
    w16 = S.pack_weight_f16(wt)
    xq = S.pack_q4t(x1)
    zoff = torch.zeros(B, dg * 18, H, W, device=cuda_dev)
    ones = torch.ones(B, dg * 9, H, W, device=cuda_dev)
    conv = torch.nn.functional.conv2d(x1, wt, None, 1, 1)
    y0 = S.dcn_tex(xq, S.pack_fields(zoff, ones, dg), w16)
    assert (y0 - conv).abs().max().item() <= 2e-3 * conv.abs().max().item()
    ioff = torch.randint(-3, 4, (B, dg * 18, H, W), generator=g).float().to(cuda_dev)
    import cdfo_b200
    gen = cdfo_b200.dcn._generic_modulated(x1, ioff, ones, wt, None, 1, 1, 1, 1, dg)
    yi = S.dcn_tex(xq, S.pack_fields(ioff, ones, dg), w16)
    assert (yi - gen).abs().max().item() <= 2e-3 * gen.abs().max().item()
    rnd = S.pack_fields((torch.randn(B, dg * 18, H, W, generator=g) * 2.0).to(cuda_dev), torch.rand(B, dg * 9, H, W, generator=g).to(cuda_dev), dg)
    off, msk = S.unpack_fields(rnd)
    gen = cdfo_b200.dcn._generic_modulated(x1, off, msk, wt, None, 1, 1, 1, 1, dg)
    yr = S.dcn_tex(xq, rnd, w16)
    assert (yr - gen).abs().max().item() <= 6e-3 * gen.abs().max().item()


@pytest.mark.parametrize("B,H,W", [(2, 37, 70), (7, 64, 96), (1, 4, 32)])
def test_dcn_tex_tma_fields_equals_ldg_fields(cuda_dev, B, H, W):
    """dg = 16, dense fields: the TMA-staged shared-memory ring and the per-thread LDG.128 path feed the same arithmetic ->
    bit-identical outputs, ragged tiles (rows / columns past the frame are zero-filled by TMA, never stored) included."""
    import cdfo_b200
    from cdfo_b200 import dcn_sm100 as S
    dg = 16
    x, offset, mask, wt, b = _case(B, H, W, dg, seed=B * H + W, off_scale=3.0)
    g = torch.Generator().manual_seed(W)
    flow = torch.randint(-192, 192, (B, 2, H, W), generator=g).float() / 128.0
    d = lambda t: t.to(cuda_dev)
    xq, fields, w16 = S.pack_q4t(d(x)), d(_fields(offset, mask, dg)), S.pack_weight_f16(d(wt))
    lib = cdfo_b200._lib.lib()
    try:
        lib.cdfo_dcn_tex_sm100_set_fields_path(1)
        y_tma = S.dcn_tex(xq, fields, w16, d(b), mv=d(flow))
        y8_tma = S.dcn_tex(xq, fields, w16, d(b), mv=d(flow), out_c8=True)
        lib.cdfo_dcn_tex_sm100_set_fields_path(0)
        y_ldg = S.dcn_tex(xq, fields, w16, d(b), mv=d(flow))
        y8_ldg = S.dcn_tex(xq, fields, w16, d(b), mv=d(flow), out_c8=True)
    finally:
        lib.cdfo_dcn_tex_sm100_set_fields_path(1)
    assert torch.equal(y_tma, y_ldg) and torch.equal(y8_tma, y8_ldg)


@pytest.mark.parametrize("xB,H,W", [(1, 40, 40), (4, 272, 40)])
def test_dcn_tex_footprint_probe_vs_oracle(cuda_dev, xB, H, W):
    """Effective sampling footprint of the BENCHMARKED kernel (dcn_tex_sm100_kernel: the texture unit receives the fp32
    coordinate base + (residual + mv), arch/SIDECVSR_our.py:3347 + deform_conv_cuda_kernel.cu:614-617, and filters with 8-bit
    fractional weights) against the exact fp32 oracle, measured with probe images instead of inferred:

      * every group's 4 channels are (local row ramp, local column ramp, one-hot anchor rows, one-hot anchor columns) -- anchors
        every 40 pixels -- and the weight is a per-tap identity, so output channel 4g+c of launch `tap` IS the blended probe value
        of sample (pixel, g, tap): ramps -> effective coordinate, one-hot -> the bilinear weight on the anchor row / column;
      * every sample is steered into [a-1, a+1) around its cell's anchor a (residuals up to +-23 px: fp16 ulp 1/64), onto exact
        integers, integers -+ 2^-17 / 2^-11 / 2^-9 (per-pixel MV perturbations the fp16 residual cannot absorb), k/128 positions
        and random fractions;
      * (4, 272, 40) is the benchmarked geometry: the texture folds (sample, quad plane, row) into its row axis, so sample 3 of
        x_batch = 4 at 272 rows samples at row coordinates up to 17 600, where fp32 keeps 9-10 fractional bits.

    Reports (and bounds) how many samples put weight on a row / column the exact footprint does not touch, the largest such weight,
    and the largest weight / effective-coordinate error."""
    from cdfo_b200 import dcn_sm100 as S
    dg = 16
    g = torch.Generator().manual_seed(12 + H)
    rows = torch.arange(H).float().view(1, 1, H, 1).expand(1, 1, H, W)
    cols = torch.arange(W).float().view(1, 1, 1, W).expand(1, 1, H, W)
    anc_r, anc_c = 40.0 * torch.floor(rows / 40.0) + 20.0, 40.0 * torch.floor(cols / 40.0) + 20.0
    ramp_r = torch.where((rows - anc_r).abs() <= 2, (rows - anc_r) * 0.5, torch.zeros(()))   # exact in fp16; linear on the footprint rows
    ramp_c = torch.where((cols - anc_c).abs() <= 2, (cols - anc_c) * 0.5, torch.zeros(()))
    quad = torch.cat([ramp_r, ramp_c, (rows == anc_r).float(), (cols == anc_c).float()], 1)    # [1, 4, H, W]
    x = quad.repeat(1, dg, 1, 1).contiguous()
    # target position of sample (pixel, group, tap): a_r - 1 + u, a_c - 1 + v with u, v in [0, 2) drawn from the hard cases
    n = dg * 9 * H * W
    kind = torch.randint(0, 6, (2, n), generator=g)
    base_int = torch.randint(0, 2, (2, n), generator=g).float()
    frac = torch.rand(2, n, generator=g)
    u = torch.where(kind <= 1, base_int, base_int + frac)                                      # 0, 1: exact integer (+ the MV delta)
    u = torch.where(kind == 2, base_int + 2.0 ** -11, u)                                       # 2: just above (near pixels only)
    u = torch.where(kind == 3, base_int + torch.randint(0, 128, (2, n), generator=g).float() / 128.0, u)   # 3: k/128 (MV grid)
    u = torch.where(kind == 4, base_int + 1.0 - 2.0 ** -9, u)                                  # 4: half a filter step below
    u = u.clamp(0.0, 2.0 - 2.0 ** -10).view(2, dg, 9, H, W)
    taps_i = torch.arange(9).view(1, 9, 1, 1) // 3
    taps_j = torch.arange(9).view(1, 9, 1, 1) % 3
    base_y = (rows[0] - 1).view(1, 1, H, W) + taps_i                                           # h*1 - 1 + i   (.cu:607-608,614)
    base_x = (cols[0] - 1).view(1, 1, H, W) + taps_j
    # MV prior on a k/64 grid (mv2mvs produces k/128; /64 keeps integer targets exactly representable as fp16 residuals up to 32 px)
    # plus a per-pixel perturbation the fp16 residual cannot absorb: integer targets become integer + delta at EVERY pixel
    mv_grid = torch.randint(-96, 96, (1, 2, H, W), generator=g).float() / 64.0
    deltas = torch.tensor([0.0, -2.0 ** -11, 2.0 ** -11, -2.0 ** -9, -2.0 ** -17, 2.0 ** -17])
    mv = mv_grid + deltas[torch.randint(0, 6, (1, 2, H, W), generator=g)]
    off_y = ((anc_r[0] - 1 + u[0]) - base_y - mv_grid[0, 1]).half().float()                    # what the fields can hold (fp16)
    off_x = ((anc_c[0] - 1 + u[1]) - base_x - mv_grid[0, 0]).half().float()
    assert float(off_y.abs().max()) <= 24 and float(off_x.abs().max()) <= 24
    offset = torch.stack([off_y, off_x], 2).reshape(1, dg * 18, H, W).contiguous()             # channel 2*(g*9+tap) + {0: dy, 1: dx}
    mask = torch.ones(1, dg * 9, H, W)
    full = offset + mv.flip(1).repeat(1, dg * 9, 1, 1)                                          # residual + flow in fp32 (arch:3347)
    d = lambda t: t.to(cuda_dev)
    rep = lambda t: t.repeat(xB, 1, 1, 1)                                                       # xB identical samples; the last one is compared
    xq, fields, mvd = S.pack_q4t(d(rep(x))), d(S.pack_fields(rep(offset), rep(mask), dg)), d(rep(mv))
    n_idx = 0
    worst_w = worst_only = worst_coord = 0.0
    eye = torch.eye(64)
    for tap in range(9):
        wt = torch.zeros(64, 64, 3, 3)
        wt[:, :, tap // 3, tap % 3] = eye
        ref = O.dcn_forward(x.numpy(), full.numpy(), mask.numpy(), wt.numpy(), None, 1, 1, 1, 1, dg)
        y = S.dcn_tex(xq, fields, S.pack_weight_f16(d(wt)), None, mv=mvd)[xB - 1].cpu().numpy()
        yq, rq = y.reshape(dg, 4, H, W), ref.reshape(dg, 4, H, W)
        # ramps: effective coordinate relative to the anchor in pixels (ramp slope 0.5)
        worst_coord = max(worst_coord, float(np.abs(yq[:, :2] - rq[:, :2]).max()) * 2.0)
        # one-hot row / column: weight on the anchor row / column; exact weights come from the oracle run itself
        worst_w = max(worst_w, float(np.abs(yq[:, 2:] - rq[:, 2:]).max()))
        # the texture footprint puts weight on a row / column the exact one does not touch = its floor index differs
        mism = (yq[:, 2:] > 0) & ~(rq[:, 2:] > 0)
        n_idx += int(mism.sum())
        if mism.any():
            worst_only = max(worst_only, float(yq[:, 2:][mism].max()))
    print("tex footprint probe xB=%d %dx%d: %d samples x 2 axes; rows / columns touched by the texture footprint only: %d (largest weight "
          "there %.3g); max |weight error| %.3g (filter step 2^-8 = %.3g); max |effective coordinate error| %.3g px"
          % (xB, H, W, n, n_idx, worst_only, worst_w, 2.0 ** -8, worst_coord))
    # the hardware quantises the fractional position to 8 bits (1.8 fixed point): weights and effective coordinates within one
    # filter step (+ fp16 rounding of the A operand, + the fp32 resolution of the folded row coordinate, 2^-10 at 16 384 rows); a
    # row / column the exact footprint does not touch never gets more than that
    bound = 2.0 ** -8 + 2.0 ** -10 + (2.0 ** -10 if xB * 16 * (H + 3) > 8192 else 0.0)
    assert worst_w <= bound and worst_only <= bound and worst_coord <= bound
