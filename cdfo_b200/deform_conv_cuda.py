"""Drop-in for the reference's compiled module `deform_conv_cuda`.

The reference binds five functions with pybind11
(ops/dcn/src/deform_conv_cuda.cpp:681-695) and calls them from
ops/dcn/deform_conv.py:52-57,76-92,144-148,161-166.  This module exports the
same five names with the same positional signatures, over the C ABI in
include/cdfo_b200.h.  Forward functions write into the caller-allocated
`output` in place; `columns` / `ones` scratch arguments are accepted and
ignored (no im2col buffer exists on this path).  Backward functions raise:
the path is inference-only.
"""
import torch

from . import _lib

__all__ = [
    "deform_conv_forward_cuda", "deform_conv_backward_input_cuda",
    "deform_conv_backward_parameters_cuda", "modulated_deform_conv_cuda_forward",
    "modulated_deform_conv_cuda_backward",
]


def _same_dtype(*ts):
    d = ts[0].dtype
    for t in ts:
        if t is not None and t.dtype != d:
            raise RuntimeError("expected all tensors to have dtype %s, got %s" % (d, t.dtype))


def deform_conv_forward_cuda(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH,
                             dilationW, dilationH, group, deformable_group, im2col_step):
    """DCNv1 forward; replaces deform_conv_forward_cuda (deform_conv_cuda.cpp:151-258). Returns 1 like it."""
    _lib.require_cuda(input, weight, offset, output)
    if weight.dim() != 4:
        raise RuntimeError("4D weight tensor (nOutputPlane,nInputPlane,kH,kW) expected, but got: %d" % weight.dim())
    if weight.size(2) != kH or weight.size(3) != kW:
        raise RuntimeError("kernel size should be consistent with weight")
    if input.dim() not in (3, 4):
        raise RuntimeError("3D or 4D input tensor expected but got: %d" % input.dim())
    # the reference makes these contiguous itself (deform_conv_cuda.cpp:166-168)
    x = input.contiguous()
    off = offset.contiguous()
    w = weight.contiguous()
    if x.dim() == 3:
        x, off = x.unsqueeze(0), off.unsqueeze(0)
    _same_dtype(x, off, w, output)
    B, C, H, W = x.shape
    if off.size(0) != B:
        raise RuntimeError("invalid batch size of offset")
    if C != w.size(1) * group:
        raise RuntimeError("invalid number of input planes, expected: %d, but got: %d" % (w.size(1) * group, C))
    if off.size(1) != deformable_group * 2 * kH * kW:
        raise RuntimeError("invalid number of channels of offset")
    Ho = (H + 2 * padH - (dilationH * (kH - 1) + 1)) // dH + 1
    Wo = (W + 2 * padW - (dilationW * (kW - 1) + 1)) // dW + 1
    if off.size(2) != Ho or off.size(3) != Wo:
        raise RuntimeError("invalid spatial size of offset, expected height: %d width: %d, but got height: %d "
                           "width: %d" % (Ho, Wo, off.size(2), off.size(3)))
    out = output if output.is_contiguous() else torch.empty_like(output, memory_format=torch.contiguous_format)
    _lib.call("cdfo_dcn_fwd", _lib.ptr(x), _lib.ptr(off), _lib.ptr(None), _lib.ptr(w), _lib.ptr(None),
                                 _lib.ptr(out), B, C, H, W, w.size(0), kH, kW, dH, dW, padH, padW,
                                 dilationH, dilationW, group, deformable_group, _lib.dtype_code(x),
                                 _lib.stream_ptr(x.device))
    if out is not output:
        output.copy_(out)
    return 1


def modulated_deform_conv_cuda_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h,
                                       kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                                       group, deformable_group, with_bias):
    """DCNv2 forward; replaces modulated_deform_conv_cuda_forward (deform_conv_cuda.cpp:486-564)."""
    _lib.require_cuda(input, weight, offset, mask, output)
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")
    B, C, H, W = input.shape
    Co, Ck, kh_, kw_ = weight.shape
    if kh_ != kernel_h or kw_ != kernel_w:
        raise RuntimeError("Input shape and kernel shape wont match: (%d x %d vs %d x %d)."
                           % (kernel_h, kernel_w, kh_, kw_))
    if C != Ck * group:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (C, Ck * group))
    off = offset.contiguous()
    msk = mask.contiguous()
    b = bias.contiguous() if with_bias else None
    _same_dtype(input, off, msk, weight, b, output)
    out = output if output.is_contiguous() else torch.empty_like(output, memory_format=torch.contiguous_format)
    _lib.call("cdfo_dcn_fwd", _lib.ptr(input), _lib.ptr(off), _lib.ptr(msk), _lib.ptr(weight), _lib.ptr(b),
                                 _lib.ptr(out), B, C, H, W, Co, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                 pad_w, dilation_h, dilation_w, group, deformable_group,
                                 _lib.dtype_code(input), _lib.stream_ptr(input.device))
    if out is not output:
        output.copy_(out)


def _inference_only(name):
    def f(*args, **kwargs):
        raise NotImplementedError(
            "%s: cdfo_b200 is an inference path; the backward of the deformable convolution is out of scope" % name)
    f.__name__ = name
    return f


deform_conv_backward_input_cuda = _inference_only("deform_conv_backward_input_cuda")
deform_conv_backward_parameters_cuda = _inference_only("deform_conv_backward_parameters_cuda")
modulated_deform_conv_cuda_backward = _inference_only("modulated_deform_conv_cuda_backward")
