"""Additional smoke checks (called by __graft_entry__.smoke): the tcgen05 / texture DCN kernel and one tiny forward of
the whole hot path (alignment, attention, fusion, trunk, tail) checked against the oracle.  The oracle is imported here
only as the checker (smoke is test infrastructure)."""


def run(dev):
    import numpy as np
    import torch
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    from cdfo_b200 import dcn_sm100 as S, synthetic
    from cdfo_b200.model import CVSR_V8
    from oracle import c_oracle as O, priors_ref, torch_ref

    # 1. texture-gather DCN at a ragged size
    g = torch.Generator().manual_seed(1)
    B, H, W, dg = 1, 20, 44, 16
    x = torch.randn(B, 64, H, W, generator=g).half().float()
    off = (torch.randn(B, dg * 18, H, W, generator=g) * 3).half().float()
    msk = torch.rand(B, dg * 9, H, W, generator=g).half().float()
    wt = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).half().float()
    b = torch.randn(64, generator=g)
    ref = O.dcn_forward(x.numpy(), off.numpy(), msk.numpy(), wt.numpy(), b.numpy(), 1, 1, 1, 1, dg)
    o = off.view(B, dg * 9, 2, H, W)
    fields = torch.stack([o[:, :, 0], o[:, :, 1], msk, torch.zeros_like(msk)], dim=-1).half().contiguous()
    y = S.dcn_tex(S.pack_q4t(x.to(dev)), fields.to(dev), S.pack_weight_f16(wt.to(dev)), b.to(dev))
    err = float(np.abs(y.cpu().numpy() - ref).max())
    assert err <= 6e-3 * float(np.abs(ref).max()), err
    print("smoke: tcgen05 + texture-gather DCN max|err| vs oracle = %.3g (max|ref| %.3g)" % (err, np.abs(ref).max()))

    # 2. one tiny forward of the DCN-alignment model (first frame) vs the torch oracle
    H, W = 24, 40
    m = CVSR_V8(alignment="mv_dcn")
    sd = synthetic.seeded_state_dict(m.state_dict(), seed=4)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    clip = synthetic.make_clip(11, H, W, 1)
    mvs = torch.from_numpy(priors_ref.mv2mvs_model_layout(clip["mv_l0"][0].numpy()))
    noise = synthetic.gumbel_uniforms(4, 1, 0, 1, H, W)
    with torch.no_grad():
        ref_sr, _ = torch_ref.cvsr_v8_forward(sd, clip["x"], mvs, clip["pms"], clip["rms"], clip["ufs"], None, noise, "O2")
    c = {k: v.to(dev) for k, v in clip.items()}
    sr, _ = m(c["x"], None, mvs.to(dev), c["pms"], c["rms"], c["ufs"], None, noise=noise)
    err = float((sr.float().cpu() - ref_sr).abs().max())
    assert err <= 1e-2, err
    print("smoke: CVSR_V8 (DCN alignment) 24x40 -> 96x160 max|err| vs oracle = %.3g" % err)
