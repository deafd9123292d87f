"""GPU parity of the hot-path modules and of the whole CVSR_V8 forward against fixtures produced by the REAL
reference (tests/golden), plus oracle comparisons at sizes the fixtures do not cover."""
import numpy as np
import pytest
import torch

import golden_util as G
from oracle import torch_ref

pytestmark = pytest.mark.gpu

# Tolerances (BASELINE.json north_star): max abs error <= 1e-2 on outputs in [0,1]; |dPSNR| <= 0.02 dB.
TOL_ABS = 1e-2
TOL_PSNR = 0.02


def _model(variant, dev, lowp=None):
    from cdfo_b200.model import CVSR_V8
    m = CVSR_V8(alignment="dual_att" if variant == "O1" else "mv_dcn")
    m.load_state_dict(G.seeded_weights(variant), strict=True)
    m = m.to(dev).eval()
    m.lowp = lowp
    return m


def _dev(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


def test_module_long_range_attention(cuda_dev):
    g = G.load("modules_golden.npz")
    d = _dev(G.module_inputs(), cuda_dev)
    m = _model("O1", cuda_dev)
    out = m.RDAB(d["res"], d["x"], d["u"])
    err = np.abs(out.cpu().numpy() - g["lra_out"]).max()
    print("RDAB max err %.3g (max|ref| %.3g)" % (err, np.abs(g["lra_out"]).max()))
    assert err <= 5e-3


def test_module_dual_att_alignment(cuda_dev):
    g = G.load("modules_golden.npz")
    d = _dev(G.module_inputs(), cuda_dev)
    m = _model("O1", cuda_dev)
    out = m.MV_deform_align(d["x"], d["extra"], d["pred"], d["flow"])
    err = np.abs(out.cpu().numpy() - g["dual_att_out"]).max()
    print("DualAttAlignment max err %.3g (max|ref| %.3g)" % (err, np.abs(g["dual_att_out"]).max()))
    assert err <= 5e-3


def test_module_mv_dcn_alignment(cuda_dev):
    from cdfo_b200 import hotpath
    g = G.load("modules_golden.npz")
    d = _dev(G.module_inputs(), cuda_dev)
    m = _model("O2", cuda_dev)
    fields = hotpath.mv_offset_fields(m.MV_deform_align, d["x"], d["extra"], d["pred"], d["flow"])
    res, msk = hotpath.unpack_fields(fields)
    off = res + d["flow"].flip(1).repeat(1, 144, 1, 1)
    e_off = np.abs(torch.cat([off[:, :18], off[:, -18:]], 1).cpu().numpy() - g["mv_offset_g0g15"]).max()
    e_msk = np.abs(torch.cat([msk[:, :9], msk[:, -9:]], 1).cpu().numpy() - g["mv_mask_g0g15"]).max()
    out = m.MV_deform_align(d["x"], d["extra"], d["pred"], d["flow"])
    err = np.abs(out.cpu().numpy() - g["mv_dcn_out"]).max()
    print("MVDualAttAlignment offset err %.3g mask err %.3g out err %.3g (max|ref| %.3g)"
          % (e_off, e_msk, err, np.abs(g["mv_dcn_out"]).max()))
    assert e_off <= 2e-2 and e_msk <= 2e-3 and err <= 5e-3


def test_module_tail(cuda_dev):
    from cdfo_b200 import hotpath
    g = G.load("modules_golden.npz")
    d = _dev(G.module_inputs(), cuda_dev)
    m = _model("O1", cuda_dev)
    out = hotpath.tail(m, d["trunk_out"], d["x_center"])
    err = np.abs(out.cpu().numpy() - g["tail_out"]).max()
    print("tail max err %.3g" % err)
    assert err <= 2e-3


@pytest.mark.parametrize("variant", ["O1", "O2"])
@pytest.mark.parametrize("lowp", [None, torch.bfloat16])
def test_full_model_vs_reference_golden(cuda_dev, variant, lowp):
    """Config c1 of BASELINE.json (7 x 64x64 LR -> 256x256): first frame and a second frame through the cache."""
    g = G.load("model_golden.npz")
    (c0, m0, n0), (c1, m1, n1) = G.two_frames()
    m = _model(variant, cuda_dev, lowp)
    c0, c1 = _dev(c0, cuda_dev), _dev(c1, cuda_dev)
    sr0, l1 = m(c0["x"], None, m0.to(cuda_dev), c0["pms"], c0["rms"], c0["ufs"], None, noise=n0)
    sr1, _ = m(c1["x"], None, m1.to(cuda_dev), c1["pms"], c1["rms"], c1["ufs"], l1, noise=n1)
    tgt_rng = np.random.default_rng(0)
    for name, sr in (("sr0", sr0), ("sr1", sr1)):
        ref = g["%s_%s" % (variant, name)]
        out = sr.float().cpu().numpy()
        err = np.abs(out - ref).max()
        target = np.clip(ref + tgt_rng.normal(0, 0.02, ref.shape), 0, 1)   # synthetic HR ground truth, ~34 dB
        dpsnr = abs(G.psnr(np.clip(out, 0, 1), target) - G.psnr(np.clip(ref, 0, 1), target))
        print("CVSR_V8 %s lowp=%s %s: max abs err %.3g, dPSNR %.4f dB" % (variant, lowp, name, err, dpsnr))
        assert err <= TOL_ABS and dpsnr <= TOL_PSNR


def test_full_model_benchmarked_config_c3_vs_reference_golden(cuda_dev):
    """bench.py's exact step against the REAL reference (tests/golden/model_c3_golden.npz, oracle/make_golden_c3.py): BASELINE.json
    configs[2] = 7 x 272x480 LR -> 1088x1920, DCN alignment (O2), B = 2 sequences, lowp = bf16, feature ring on, the six noise
    tensors handed over as one neighbour-major batch, texture-gather DCN -- a first frame and a second frame through the ring."""
    import cdfo_b200
    g = G.load("model_c3_golden.npz")
    (c0, m0, n0), (c1, m1, n1) = G.c3_frames()
    assert cdfo_b200.config.dcn_gather == "tex" and cdfo_b200.config.head_dual
    m = _model("O2", cuda_dev, torch.bfloat16)
    m.feature_ring = True
    c0, c1 = _dev(c0, cuda_dev), _dev(c1, cuda_dev)
    sr0, l1 = m(c0["x"], None, m0.to(cuda_dev), c0["pms"], c0["rms"], c0["ufs"], None, noise=torch.cat(n0, 0).to(cuda_dev))
    assert getattr(l1, "_cdfo_ring", None) is not None
    sr1, _ = m(c1["x"], None, m1.to(cuda_dev), c1["pms"], c1["rms"], c1["ufs"], l1, noise=torch.cat(n1, 0).to(cuda_dev))
    for i, sr in enumerate((sr0, sr1)):
        ref16 = g["sr%d" % i]
        out = sr.float().cpu().numpy()
        err = np.abs(out - ref16.astype(np.float32)).max()          # includes <= 2.5e-4 of fp16 storage of the golden
        target = G.c3_target(ref16, i)
        dpsnr = np.abs(G.psnr_per_sequence(np.clip(out, 0, 1), target) - g["psnr%d" % i])
        print("CVSR_V8 O2 c3 272x480 B=2 bf16 ring %s frame: max abs err %.3g, dPSNR per sequence %s dB (reference PSNR %s)"
              % (("first", "cached")[i], err, np.round(dpsnr, 5), np.round(g["psnr%d" % i], 3)))
        assert err <= TOL_ABS and dpsnr.max() <= TOL_PSNR


def test_full_model_ragged_size_vs_oracle(cuda_dev):
    """A size with a ragged last DCN tile (W = 40, not a multiple of 32), B = 2, against the torch oracle."""
    from cdfo_b200 import synthetic
    H, W, B = 24, 40, 2
    clip = synthetic.make_clip(7, H, W, B)
    from oracle import priors_ref
    mvs = torch.stack([torch.from_numpy(priors_ref.mv2mvs_model_layout(clip["mv_l0"][b].numpy())[0]) for b in range(B)])
    noise = synthetic.gumbel_uniforms(4, 3, 0, B, H, W)
    sd = G.seeded_weights("O2")
    with torch.no_grad():
        ref, _ = torch_ref.cvsr_v8_forward(sd, clip["x"], mvs, clip["pms"], clip["rms"], clip["ufs"], None, noise, "O2")
    m = _model("O2", cuda_dev)
    c = _dev(clip, cuda_dev)
    sr, _ = m(c["x"], None, mvs.to(cuda_dev), c["pms"], c["rms"], c["ufs"], None, noise=noise)
    err = (sr.cpu() - ref).abs().max().item()
    print("ragged O2 B=2 24x40: max abs err %.3g" % err)
    assert err <= TOL_ABS


def test_full_model_config_c2_vs_oracle(cuda_dev):
    """BASELINE.json configs[1] (JCT-VC Class-C-shaped clip: 7 x 208x120 LR -> 832x480 HR, LD priors), first frame and a cached
    second frame, bf16 trunk / feature extraction, against the fp32 torch oracle on the host CPU (~10 s per frame)."""
    from cdfo_b200 import synthetic
    from oracle import priors_ref
    H, W = 120, 208
    sd = G.seeded_weights("O2")
    m = _model("O2", cuda_dev, torch.bfloat16)
    clip = synthetic.make_clip(21, H, W, 1)
    mvs = torch.from_numpy(priors_ref.mv2mvs_model_layout(clip["mv_l0"][0].numpy()))
    n0, n1 = synthetic.gumbel_uniforms(4, 2, 0, 1, H, W), synthetic.gumbel_uniforms(4, 2, 1, 1, H, W)
    with torch.no_grad():
        ref0, l1_ref = torch_ref.cvsr_v8_forward(sd, clip["x"], mvs, clip["pms"], clip["rms"], clip["ufs"], None, n0, "O2")
        ref1, _ = torch_ref.cvsr_v8_forward(sd, clip["x"], mvs, clip["pms"], clip["rms"], clip["ufs"], l1_ref, n1, "O2")
    c = _dev(clip, cuda_dev)
    sr0, l1 = m(c["x"], None, mvs.to(cuda_dev), c["pms"], c["rms"], c["ufs"], None, noise=n0)
    sr1, _ = m(c["x"], None, mvs.to(cuda_dev), c["pms"], c["rms"], c["ufs"], l1, noise=n1)
    rng = np.random.default_rng(1)
    for name, sr, ref in (("first", sr0, ref0), ("cached", sr1, ref1)):
        out, r = sr.float().cpu().numpy(), ref.numpy()
        err = np.abs(out - r).max()
        target = np.clip(r + rng.normal(0, 0.02, r.shape), 0, 1)
        dpsnr = abs(G.psnr(np.clip(out, 0, 1), target) - G.psnr(np.clip(r, 0, 1), target))
        print("CVSR_V8 O2 c2 120x208 %s frame: max abs err %.3g, dPSNR %.4f dB" % (name, err, dpsnr))
        assert err <= TOL_ABS and dpsnr <= TOL_PSNR


def test_full_model_config_c5_ra_priors_vs_oracle(cuda_dev):
    """BASELINE.json configs[4] in miniature: the same network fed RA priors -- (a) the bidirectional decoding of an (l0, l1) MV
    pair with -99 sentinels (opt/data_RA_bi.py:496-533, what train_RA_37 feeds the model) and (b) the RA eval loop's choice, the
    LD-style mv2mvs applied to l1 alone (mvs1 is the only flow the forward consumes, arch:4445).  Flows must be bit-exact
    against the oracle's decoding; the SR frame within the north-star tolerance of the fp32 oracle forward."""
    import cdfo_b200
    from cdfo_b200 import synthetic
    from oracle import priors_ref
    H, W = 40, 64
    clip = synthetic.make_clip(31, H, W, 1, config="RA")
    l0, l1 = clip["mv_l0"][0], clip["mv_l1"][0]
    assert int((l1[..., 2] == -99).sum()) > 0
    sd = G.seeded_weights("O2")
    m = _model("O2", cuda_dev)
    c = _dev({k: clip[k] for k in ("x", "pms", "rms", "ufs")}, cuda_dev)
    with np.errstate(all="ignore"):
        flows_ra = priors_ref.mv2mvs_ra(l0.numpy(), l1.numpy())                                  # [7, H, W, 2]
        flows_l1 = priors_ref.mv2mvs(l1.numpy())
    cases = {"bidirectional": (np.ascontiguousarray(flows_ra.transpose(0, 3, 1, 2))[None], cdfo_b200.mv2mvs_ra(l0.to(cuda_dev), l1.to(cuda_dev))),
             "eval-l1": (np.ascontiguousarray(np.asarray(flows_l1).transpose(0, 3, 1, 2))[None], cdfo_b200.mv2mvs(l1.to(cuda_dev)))}
    for k, (ref_flows, dev_flows) in cases.items():
        assert np.array_equal(dev_flows.cpu().numpy().view(np.uint32), ref_flows.view(np.uint32)), k     # bit-exact, inf included
        if not np.isfinite(ref_flows).all():      # x / 0 stays +-inf in the reference (test_LD_37.py:93 only filters NaN)
            continue
        noise = synthetic.gumbel_uniforms(4, 5, 0, 1, H, W)
        with torch.no_grad():
            ref, _ = torch_ref.cvsr_v8_forward(sd, clip["x"], torch.from_numpy(ref_flows), clip["pms"], clip["rms"], clip["ufs"], None, noise, "O2")
        sr, _ = m(c["x"], None, dev_flows, c["pms"], c["rms"], c["ufs"], None, noise=noise)
        err = (sr.cpu() - ref).abs().max().item()
        print("CVSR_V8 O2 RA priors (%s) 40x64: max abs err %.3g" % (k, err))
        assert err <= TOL_ABS


def test_graphed_step_matches_eager(cuda_dev):
    """cdfo_b200.graph.GraphedStep: the steady-state step captured in a CUDA graph replays to bit-identical outputs."""
    from cdfo_b200 import synthetic
    from cdfo_b200.graph import GraphedStep
    from oracle import priors_ref
    H, W, B = 32, 48, 2
    m = _model("O2", cuda_dev, torch.bfloat16)
    clip = synthetic.make_clip(5, H, W, B)
    c = _dev(clip, cuda_dev)
    mvs = torch.stack([torch.from_numpy(priors_ref.mv2mvs_model_layout(clip["mv_l0"][b].numpy())[0]) for b in range(B)]).to(cuda_dev)
    noise = [u.to(cuda_dev) for u in synthetic.gumbel_uniforms(4, 9, 0, B, H, W)]
    _, l1 = m(c["x"], None, mvs, c["pms"], c["rms"], c["ufs"], None, noise=noise)
    sr_e, l1_e = m(c["x"], None, mvs, c["pms"], c["rms"], c["ufs"], l1, noise=noise)
    gs = GraphedStep(m, c["x"], mvs, c["pms"], c["rms"], c["ufs"], l1, noise)
    for _ in range(2):
        sr_g, l1_g = gs(c["x"], mvs, c["pms"], c["rms"], c["ufs"], l1, noise)
        torch.cuda.synchronize()
        assert torch.equal(sr_g, sr_e) and torch.equal(l1_g, l1_e)


def test_feature_ring_matches_window_cat(cuda_dev):
    """model.feature_ring: the L1 features of the sliding window kept as a frame-major ring (no per-frame copy of the window, RDAB on
    the two contiguous neighbour runs, noise passed pre-concatenated) gives the outputs of the plain `pre_L1_fea` path over
    several steps; a stale handle is refused."""
    from cdfo_b200 import synthetic, _lib
    from oracle import priors_ref
    H, W, B = 32, 48, 2
    m = _model("O2", cuda_dev, torch.bfloat16)
    clip = synthetic.make_clip(7, H, W, B)
    c = _dev(clip, cuda_dev)
    mvs = torch.stack([torch.from_numpy(priors_ref.mv2mvs_model_layout(clip["mv_l0"][b].numpy())[0]) for b in range(B)]).to(cuda_dev)
    noise = [u.to(cuda_dev) for u in synthetic.gumbel_uniforms(4, 9, 0, B, H, W)]

    def run(ring, steps=9):      # more steps than ring slots: the head wraps
        m.feature_ring = ring
        srs, l1, first = [], None, None
        for s in range(steps):
            x = torch.roll(c["x"], s, dims=1)           # a different "new frame" every step
            pms = torch.roll(c["pms"], s, dims=1)
            sr, l1 = m(x, None, mvs, pms, c["rms"], c["ufs"], l1, noise=torch.cat(noise, 0) if ring else noise)
            first = l1 if first is None else first
            srs.append(sr)
        m.feature_ring = False
        return srs, l1, first

    srs_a, l1_a, _ = run(False)
    srs_b, l1_b, stale = run(True)
    for s, (a, b) in enumerate(zip(srs_a, srs_b)):
        err = (a - b).abs().max().item()
        assert err <= 2e-3, "step %d: ring vs cat differ by %.3g" % (s, err)      # cuDNN picks algorithms per call
    assert torch.equal(l1_b.view(7, B, 64, H, W).transpose(0, 1).reshape(B * 7, 64, H, W), l1_a)
    with pytest.raises(_lib.CdfoError):
        m(c["x"], None, mvs, c["pms"], c["rms"], c["ufs"], stale, noise=noise)


def test_dual_head_launch_equals_two_launches(cuda_dev):
    """cdfo_mv_offset_head_dual_sm100_fwd (both evaluations of conv_offset[-1] in one launch, the first kept in registers)
    writes the same fields, bit for bit, as the two-launch path with its fp16 intermediate in HBM (ragged tiles included)."""
    from cdfo_b200 import config, hotpath
    m = _model("O2", cuda_dev)
    g = torch.Generator().manual_seed(21)
    B, H, W = 3, 40, 56                      # 40 x 56: partial 16 x 8 tiles on both axes
    x = torch.randn(1, 64, H, W, generator=g).to(cuda_dev)
    extra, pred = torch.randn(B, 64, H, W, generator=g).to(cuda_dev), torch.randn(B, 64, H, W, generator=g).to(cuda_dev)
    flow = (torch.randn(B, 2, H, W, generator=g) * 2).to(cuda_dev)
    out = {}
    for dual in (True, False):
        config.head_dual = dual
        out[dual] = hotpath.mv_offset_fields(m.MV_deform_align, x, extra, pred, flow)
    config.head_dual = True
    assert torch.equal(out[True].view(torch.int16), out[False].view(torch.int16))
