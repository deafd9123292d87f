"""Per-stage CUDA-event timing of one steady-state frame (c3, S sequences) -- where the frame time goes."""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdfo_b200  # noqa: E402
from cdfo_b200 import hotpath, synthetic  # noqa: E402
from cdfo_b200.model import CVSR_V8  # noqa: E402


def t(fn, iters=5, warmup=2):
    for _ in range(warmup):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=2)
    ap.add_argument("--H", type=int, default=272)
    ap.add_argument("--W", type=int, default=480)
    ap.add_argument("--variant", default="O2")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    S, H, W = a.seqs, a.H, a.W
    m = CVSR_V8(alignment="mv_dcn" if a.variant == "O2" else "dual_att")
    m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), 4))
    m = m.to(dev).eval()
    m.lowp = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
    x1 = torch.rand(S, 1, H, W, device=dev, generator=g)
    fea_nb, center = r(6 * S, 64, H, W), r(S, 64, H, W)
    ufs_nb, rms_nb = torch.rand(6 * S, 1, H, W, device=dev, generator=g), r(6 * S, 1, H, W) * 0.02
    mv = (torch.randint(-192, 192, (6 * S, 2, H // 8, W // 8), device=dev, generator=g).float() / 128).repeat_interleave(8, 2).repeat_interleave(8, 3)
    u = torch.rand(6 * S, 64, H, W, device=dev, generator=g).clamp_min(1e-12)
    res = {}
    with torch.no_grad():
        res["features_1frame"], _ = t(lambda: m._features(x1, x1))
        res["prior_convs"], (up, rp) = t(lambda: (hotpath.prior_conv(m.conv_expand_ufs, ufs_nb), hotpath.prior_conv(m.conv_expand_rms, rms_nb)))
        res["RDAB"], x_n = t(lambda: hotpath.long_range_attention(m.RDAB, rp, fea_nb + rp, u))
        fr = m.conv_expand_fea_r
        res["conv_expand_fea_r"], fea_i = t(lambda: cdfo_b200.conv.conv3x3(cdfo_b200.conv.to_c8(torch.cat([fea_nb, x_n], 1)), fr.weight, fr.bias, 0, out_nchw=True))
        cr = center.repeat(6, 1, 1, 1)
        al = m.MV_deform_align
        res["align.dual_mdta"], _ = t(lambda: hotpath.dual_mdta(al, center, fea_i, up, mv, 0 if a.variant == "O2" else 1))
        if a.variant == "O2":
            res["align.offset_fields_total"], fields = t(lambda: hotpath.mv_offset_fields(al, center, fea_i, up, mv))
            res["align.pack_q4t"], xq = t(lambda: cdfo_b200.dcn_sm100.pack_q4t(center))
            res["align.dcn_tex"], aligned = t(lambda: cdfo_b200.dcn_sm100.dcn_tex(xq, fields, cdfo_b200.dcn_sm100.pack_weight_f16(al.weight), al.bias, mv=mv))
        res["align.total"], aligned = t(lambda: al(center, fea_i, up, mv))
        res["align_and_fuse_total"], fused = t(lambda: hotpath.align_and_fuse(m, center, fea_nb, ufs_nb, rms_nb, mv, u, S))
        res["trunk"], tr = t(lambda: m._trunk(fused))
        res["tail"], _ = t(lambda: hotpath.tail(m, tr, x1))
    res = {k: round(v, 3) for k, v in res.items()}
    res["sum_ms"] = round(res["features_1frame"] + res["align_and_fuse_total"] + res["trunk"] + res["tail"], 2)
    res["seqs"] = S
    print(json.dumps(res))


if __name__ == "__main__":
    main()
