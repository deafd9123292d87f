"""Micro-benchmark of MVDualAttAlignment's offset path at BASELINE config c3 (272x480 LR): the fused head + DCN kernel
(cdfo_mv_head_dcn_fused_sm100_fwd) against the two launches it replaces (dual head -> fp16 fields in HBM -> texture-gather DCN).
CUDA events on the launching stream; B samples = neighbour calls (6 per sequence and frame), x shared by x_batch = B / 6 sequences.
Roofline: 1 069 056 FLOP per LR pixel and call, tensor-bound; peak from MEASURED_PEAKS.json."""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cdfo_b200  # noqa: E402
from cdfo_b200 import _lib, conv, dcn_sm100 as S, hotpath  # noqa: E402

FLOP_PER_PX = 2 * 497664 + 73728


def timeit(fn, iters, warmup=3):
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
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only-fused", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, H, W = a.B, a.H, a.W
    xB = max(1, B // 6)
    g = torch.Generator(device=dev).manual_seed(0)
    z8 = conv.to_c8(torch.randn(2 * B, 64, H, W, device=dev, generator=g) * 0.5)
    c2 = torch.nn.Conv2d(64, 432, 3, 1, 1).to(dev)
    with torch.no_grad():
        c2.weight.copy_(torch.randn(432, 64, 3, 3, device=dev, generator=g) * 0.01)
        c2.bias.copy_(torch.randn(432, device=dev, generator=g) * 0.1)
    xq = S.pack_q4t(torch.randn(xB, 64, H, W, device=dev, generator=g))
    mv = torch.nn.functional.interpolate(torch.randint(-192, 192, (B, 2, H // 8, W // 8), device=dev, generator=g).float() / 128.0,
                                         scale_factor=8, mode="nearest").contiguous()
    wd16 = S.pack_weight_f16(torch.randn(64, 64, 3, 3, device=dev, generator=g) * 0.05)
    bd = torch.randn(64, device=dev, generator=g)
    stack = torch.empty((xB, 56, H, W, 8), dtype=torch.bfloat16, device=dev)
    chunks = [0, 8, 16, 32, 40, 48][:B // xB]
    hw, hb = hotpath._head_weights_fused(c2, 16)
    wpk, bias = hotpath._head_weights(c2, 16)
    fields = torch.empty(S.fields_shape(B, 16, H, W), dtype=torch.float16, device=dev)
    P = H * W
    res = {"B": B, "H": H, "W": W, "x_batch": xB, "flop_per_px": FLOP_PER_PX}

    def fused():
        S.mv_head_dcn_fused(z8, hw, hb, 10.0, xq, mv, wd16, bd, stack=stack, group_chunk=chunks)

    def head():
        _lib.call("cdfo_mv_offset_head_dual_sm100_fwd", _lib.ptr(z8), _lib.ptr(wpk), _lib.ptr(bias), _lib.ptr(fields), B, 64, 16, H, W,
                  ctypes.c_float(10.0), _lib.stream_ptr(dev))

    def dcn():
        S.dcn_tex_stacked(xq, fields, wd16, bd, mv, stack, chunks)

    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    t = timeit(fused, a.iters)
    res["fused_us"] = t
    res["fused_us_per_call"] = t / B
    res["fused_tflops"] = FLOP_PER_PX * P * B / t / 1e6
    if peaks:
        res["fused_frac_of_bf16_burst"] = res["fused_tflops"] / peaks["bf16_tflops"]
        res["fused_frac_of_bf16_sustained"] = res["fused_tflops"] / peaks["bf16_tflops_sustained"]
    if not a.only_fused:
        th, td = timeit(head, a.iters), None
        head()
        td = timeit(dcn, a.iters)
        res["head_dual_us"], res["dcn_tex_us"] = th, td
        res["two_kernel_us_per_call"] = (th + td) / B
        res["two_kernel_tflops"] = FLOP_PER_PX * P * B / (th + td) / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
