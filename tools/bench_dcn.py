"""Micro-benchmark of the DCN kernels at BASELINE config c3 (272x480 LR), B = 6 neighbour calls of one frame.
Times with CUDA events on the launching stream; inputs (>= 450 MB of offsets/masks) exceed the 126 MB L2."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdfo_b200  # noqa: E402
from cdfo_b200 import dcn_sm100 as S  # noqa: E402


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=6)
    ap.add_argument("--H", type=int, default=272)
    ap.add_argument("--W", type=int, default=480)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--smooth", type=int, default=1, help="block-constant MV-like offsets (1) or i.i.d. random (0)")
    ap.add_argument("--skip-baselines", action="store_true")
    ap.add_argument("--only-tex", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, H, W, dg = a.B, a.H, a.W, 16
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, 64, H, W, device=dev, generator=g)
    if a.smooth:
        mvb = torch.randint(-192, 192, (B, 2, H // 8, W // 8), device=dev, generator=g).float() / 128.0
        mv = torch.nn.functional.interpolate(mvb, scale_factor=8, mode="nearest")
        offset = torch.randn(B, dg * 18, H, W, device=dev, generator=g) * 0.3 + mv.flip(1).repeat(1, dg * 9, 1, 1)
    else:
        offset = torch.randn(B, dg * 18, H, W, device=dev, generator=g) * 4.0
    mask = torch.rand(B, dg * 9, H, W, device=dev, generator=g)
    wt = torch.randn(64, 64, 3, 3, device=dev, generator=g) * 0.05
    bias = torch.randn(64, device=dev, generator=g)
    xc = S.pack_q4p(x)
    wpk = S.pack_weight(wt)
    off16, m16 = offset.half(), mask.half()
    P = H * W
    res = {"B": B, "H": H, "W": W, "smooth": a.smooth}
    if not a.only_tex:
        t = timeit(lambda: S.dcn_sm100(xc, offset, mask, wpk, bias, num_ctas=a.ctas), a.iters)
        res["sm100_f32off_us"] = t
        res["sm100_f32off_GBs_fp32io_2240"] = 2240.0 * P * B / t / 1e3
        t = timeit(lambda: S.dcn_sm100(xc, off16, m16, wpk, bias, out_c8=True, num_ctas=a.ctas), a.iters)
        res["sm100_f16off_c8out_us"] = t
        res["sm100_f16off_GBs_algo_1120"] = 1120.0 * P * B / t / 1e3
        res["pack_q4p_us"] = timeit(lambda: S.pack_q4p(x), a.iters)
    fields = S.pack_fields(offset, mask, dg)
    xt, w16 = S.pack_q4t(x), S.pack_weight_f16(wt)
    t = timeit(lambda: S.dcn_tex(xt, fields, w16, bias, out_c8=True, num_ctas=a.ctas), a.iters)
    res["tex_fields_c8out_us"] = t
    res["tex_GBs_algo_1120"] = 1120.0 * P * B / t / 1e3
    if not a.only_tex:
        y_tma = S.dcn_tex(xt, fields, w16, bias)
        cdfo_b200._lib.lib().cdfo_dcn_tex_sm100_set_fields_path(0)      # A/B: fields through per-thread LDG.128 instead of TMA
        res["tex_fields_ldg_c8out_us"] = timeit(lambda: S.dcn_tex(xt, fields, w16, bias, out_c8=True, num_ctas=a.ctas), a.iters)
        res["tex_tma_vs_ldg_fields_maxdiff"] = float((y_tma - S.dcn_tex(xt, fields, w16, bias)).abs().max())
        cdfo_b200._lib.lib().cdfo_dcn_tex_sm100_set_fields_path(1)
    t = timeit(lambda: S.dcn_tex(xt, fields, w16, bias, num_ctas=a.ctas), a.iters)
    res["tex_fields_f32out_us"] = t
    res["pack_q4t_us"] = timeit(lambda: S.pack_q4t(x), a.iters)
    ya, yb = S.dcn_tex(xt, fields, w16, bias), S.dcn_sm100(xc, off16, m16, wpk, bias)
    res["tex_vs_exact_maxdiff"] = float((ya - yb).abs().max())
    res["out_absmax"] = float(yb.abs().max())
    if not a.skip_baselines:
        cdfo_b200.config.tensor_core = False
        res["generic_fp32_us"] = timeit(
            lambda: cdfo_b200.modulated_deform_conv(x, offset, mask, wt, bias, 1, 1, 1, 1, dg), 5, 1)
        cdfo_b200.config.tensor_core = True
        try:
            import torchvision
            res["torchvision_cuda_fp32_us"] = timeit(
                lambda: torchvision.ops.deform_conv2d(x, offset, wt, bias, 1, 1, 1, mask), 5, 1)
        except Exception as e:  # noqa: BLE001
            res["torchvision_cuda_fp32_us"] = "unavailable: %s" % e
        # the reference's OWN CUDA op (ops/dcn/src/*, built by baseline/build_ref_dcn.py into baseline/_ref/): the same-box figure to beat
        import glob
        import importlib.util
        so = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "deform_conv_cuda*.so")))
        if so:
            spec = importlib.util.spec_from_file_location("deform_conv_cuda", so[0])
            ref = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref)

            def ref_call(xx, oo, mm, ww, bb):
                out = xx.new_empty((B, 64, H, W))
                ref.modulated_deform_conv_cuda_forward(xx, ww, bb, xx.new_empty(0), oo, mm, out, xx.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, True)
                return out
            res["reference_ext_cuda_fp32_us"] = timeit(lambda: ref_call(x, offset, mask, wt, bias), 5, 1)
            res["reference_ext_cuda_fp16_us"] = timeit(lambda: ref_call(x.half(), off16, m16, wt.half(), bias.half()), 5, 1)
            res["reference_ext_vs_exact_maxdiff"] = float((ref_call(x, offset, mask, wt, bias) - yb).abs().max())
        else:
            res["reference_ext_cuda_fp32_us"] = "baseline/_ref not built"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
