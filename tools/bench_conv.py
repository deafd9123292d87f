"""Micro-benchmark: tcgen05 3x3 conv vs cuDNN (bf16 channels_last) at the model's shapes (c3 = 272x480)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdfo_b200 import conv  # noqa: E402
import cdfo_b200  # noqa: E402

cdfo_b200_lib = cdfo_b200._lib.lib()


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    dev = torch.device("cuda:0")
    H, W = 272, 480
    out = []
    for (B, Cin, Cout, h, w) in [(12, 64, 64, H, W), (12, 64, 432, H, W), (12, 128, 64, H, W), (2, 64, 256, H, W),
                                 (2, 256, 64, H, W), (2, 64, 256, 2 * H, 2 * W), (2, 256, 64, 2 * H, 2 * W)]:
        x = torch.randn(B, Cin, h, w, device=dev)
        wt = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
        b = torch.randn(Cout, device=dev)
        x8 = conv.to_c8(x)
        t_ours = timeit(lambda: conv.conv3x3(x8, wt, b, conv.ACT_LRELU))
        t_single = None
        if cdfo_b200_lib.cdfo_conv3x3_pair_sm100_supported(Cout, Cin):      # A/B: the single-SM kernel
            cdfo_b200.config.conv_pair = False
            t_single = timeit(lambda: conv.conv3x3(x8, wt, b, conv.ACT_LRELU))
            cdfo_b200.config.conv_pair = True
        xcl = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wcl = wt.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        bb = b.to(torch.bfloat16)
        t_cudnn = timeit(lambda: F.leaky_relu(F.conv2d(xcl, wcl, bb, 1, 1), 0.1))
        flops = 2.0 * B * h * w * Cin * Cout * 9
        out.append({"shape": [B, Cin, Cout, h, w], "ours_us": round(t_ours, 1), "ours_TFLOPs": round(flops / t_ours / 1e6, 1),
                    "single_sm_TFLOPs": None if t_single is None else round(flops / t_single / 1e6, 1),
                    "cudnn_bf16_us": round(t_cudnn, 1), "cudnn_TFLOPs": round(flops / t_cudnn / 1e6, 1)})
    # the folded "3x3 at 2x + bilinear x0.5" (4x4 stride 2 on a CTA pair) against its two-kernel form
    for B in (2, 4):
        x = torch.randn(B, 256, 2 * H, 2 * W, device=dev)
        wt = torch.randn(64, 256, 3, 3, device=dev) * 0.05
        b = torch.randn(64, device=dev)
        x8 = conv.to_c8(x)
        base = conv.to_c8(torch.randn(B, 64, H, W, device=dev))
        t_fold = timeit(lambda: conv.conv3x3_then_half(x8, wt, b, base))
        xp = x8.view(B, 32, H, 2, W, 2, 8).permute(0, 1, 3, 5, 2, 4, 6).contiguous()
        t_planes = timeit(lambda: conv.conv3x3_then_half(xp, wt, b, base))
        t_two = timeit(lambda: conv.resample(conv.conv3x3(x8, wt, b, conv.ACT_NONE), 0))
        flops = 2.0 * B * H * W * 256 * 64 * 16
        out.append({"shape": [B, 256, 64, 2 * H, 2 * W], "folded_4x4s2_us": round(t_fold, 1), "folded_TFLOPs": round(flops / t_fold / 1e6, 1),
                    "folded_parity_planes_us": round(t_planes, 1), "folded_parity_planes_TFLOPs": round(flops / t_planes / 1e6, 1),
                    "conv3x3_plus_resample_us": round(t_two, 1)})
    # 64 -> 256 at 2x writing plain c8 against parity planes
    x8 = conv.to_c8(torch.randn(2, 64, 2 * H, 2 * W, device=dev))
    wt = torch.randn(256, 64, 3, 3, device=dev) * 0.05
    b = torch.randn(256, device=dev)
    out.append({"shape": [2, 64, 256, 2 * H, 2 * W], "c8_out_us": round(timeit(lambda: conv.conv3x3(x8, wt, b, conv.ACT_LRELU)), 1),
                "parity_planes_out_us": round(timeit(lambda: conv.conv3x3(x8, wt, b, conv.ACT_LRELU, parity_planes=True)), 1)})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
