"""Host-side mirror of the reference's `ops/dcn/deform_conv.py` operator interface.

Same public names, constructor arguments, parameter names/shapes (so reference
state_dicts load) and error behaviour as ops/dcn/deform_conv.py:186-337, over
`cdfo_b200.deform_conv_cuda` (the drop-in for the compiled module).  Inference
only: tensors that require grad are accepted but no graph is recorded.

Also exports `deform_conv2d`, argument-compatible with
`torchvision.ops.deform_conv2d` -- the call the model's DCN alignment makes
(arch/SIDECVSR_our.py:3352).
"""
import math

import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from . import config, dcn_sm100, deform_conv_cuda

__all__ = ["deform_conv", "modulated_deform_conv", "deform_conv2d", "DeformConv", "DeformConvPack",
           "ModulatedDeformConv", "ModulatedDeformConvPack"]


def _out_hw(in_hw, k_hw, stride, padding, dilation):
    out = []
    for d in range(2):
        eff = dilation[d] * (k_hw[d] - 1) + 1
        out.append((in_hw[d] + 2 * padding[d] - eff) // stride[d] + 1)
    return out


@torch.no_grad()
def deform_conv(input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1,
                im2col_step=64):
    """DCNv1 forward. Mirrors DeformConvFunction.forward (ops/dcn/deform_conv.py:16-58)."""
    if input is not None and input.dim() != 4:
        raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
    stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
    oh, ow = _out_hw(input.shape[2:], weight.shape[2:], stride, padding, dilation)
    shape = (input.size(0), weight.size(0), oh, ow)
    if not all(s > 0 for s in shape):
        raise ValueError("convolution input is too small (output would be {})".format("x".join(map(str, shape))))
    if not input.is_cuda:
        raise NotImplementedError
    output = input.new_empty(shape)
    step = min(im2col_step, input.shape[0])
    assert input.shape[0] % step == 0, "im2col step must divide batchsize"
    scratch = input.new_empty(0)
    deform_conv_cuda.deform_conv_forward_cuda(
        input, weight, offset, output, scratch, scratch, weight.size(3), weight.size(2), stride[1], stride[0],
        padding[1], padding[0], dilation[1], dilation[0], groups, deformable_groups, step)
    return output


@torch.no_grad()
def _generic_modulated(input, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups):
    """fp32-exact catch-all kernel through the drop-in module (caller-side output allocation like the reference)."""
    with_bias = bias is not None
    kh, kw = weight.shape[2:4]
    oh = (input.size(2) + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    ow = (input.size(3) + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    output = input.new_empty((input.size(0), weight.size(0), oh, ow))
    scratch = input.new_empty(0)
    deform_conv_cuda.modulated_deform_conv_cuda_forward(
        input, weight, bias if with_bias else input.new_empty(1), scratch, offset, mask, output, scratch,
        kh, kw, stride, stride, padding, padding, dilation, dilation, groups, deformable_groups, with_bias)
    return output


@torch.no_grad()
def _tensor_core_modulated(input, offset, mask, weight, bias, deformable_groups, mv=None):
    """Hot shape: pack x (bf16, zero border), cached bf16 weights, one tcgen05 implicit-GEMM kernel."""
    y = dcn_sm100.dcn_sm100(dcn_sm100.pack_q4p(input), offset.float(), mask.float(), dcn_sm100.pack_weight(weight),
                            bias, mv=mv)
    return y if input.dtype == torch.float32 else y.to(input.dtype)


@torch.no_grad()
def modulated_deform_conv(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                          deformable_groups=1):
    """DCNv2 forward. Mirrors ModulatedDeformConvFunction.forward (ops/dcn/deform_conv.py:116-149):
    scalar stride/padding/dilation, CUDA only.  The model's hot shape runs on the tensor cores."""
    if not input.is_cuda:
        raise NotImplementedError
    if config.tensor_core and input.is_contiguous() and dcn_sm100.supported(
            input, weight, _pair(stride), _pair(padding), _pair(dilation), groups, deformable_groups, mask):
        return _tensor_core_modulated(input, offset, mask, weight, bias, deformable_groups)
    return _generic_modulated(input, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups)


@torch.no_grad()
def deform_conv2d(input, offset, weight, bias=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1), mask=None):
    """Argument-compatible with torchvision.ops.deform_conv2d (groups are inferred from shapes the same way)."""
    from . import _lib
    _lib.require_cuda(input, offset, weight)
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    Co, Ck, kh, kw = weight.shape
    B, C, H, W = input.shape
    dg = offset.shape[1] // (2 * kh * kw)
    groups = C // Ck
    if dg == 0:
        raise RuntimeError("the shape of the offset tensor at dimension 1 is not valid. It should be a multiple of "
                           "2 * weight.size[2] * weight.size[3].")
    oh, ow = _out_hw((H, W), (kh, kw), (sh, sw), (ph, pw), (dh, dw))
    # shape errors as torchvision raises them (torchvision/csrc/ops/cuda/deform_conv2d_kernel.cu checks)
    if groups == 0 or C != Ck * groups or Co % groups:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (C, Ck * max(groups, 1)))
    if offset.shape[1] != dg * 2 * kh * kw or C % dg:
        raise RuntimeError("offset.shape[1] is not valid: got: %d expected: %d" % (offset.shape[1], dg * 2 * kh * kw))
    if tuple(offset.shape) != (B, dg * 2 * kh * kw, oh, ow):
        raise RuntimeError("offset output dims: (%d, %d) - computed output dims: (%d, %d)" % (offset.shape[2], offset.shape[3], oh, ow))
    if mask is not None and tuple(mask.shape) != (B, dg * kh * kw, oh, ow):
        raise RuntimeError("mask.shape is not valid: got %s expected %s" % (tuple(mask.shape), (B, dg * kh * kw, oh, ow)))
    if bias is not None and tuple(bias.shape) != (Co,):
        raise RuntimeError("invalid bias shape: got: %s expected: (%d,)" % (tuple(bias.shape), Co))
    x = input.contiguous()
    # one element width for every pointer of the C ABI (cdfo_dcn_fwd takes a single dtype code): under autocast x may be fp16 / bf16
    # while offsets and weights are still fp32 -- cast them to x's dtype like torchvision's autocast wrapper does
    cast = lambda t: None if t is None else t.to(x.dtype).contiguous()  # noqa: E731
    offset, mask, weight, bias = cast(offset), cast(mask), cast(weight), cast(bias)
    if config.tensor_core and dcn_sm100.supported(x, weight, (sh, sw), (ph, pw), (dh, dw), groups, dg, mask):
        return _tensor_core_modulated(x, offset, mask, weight, bias, dg)
    y = x.new_empty((B, Co, oh, ow))
    _lib.call("cdfo_dcn_fwd",
        _lib.ptr(x), _lib.ptr(offset), _lib.ptr(mask),
        _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(y),
        B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, _lib.dtype_code(x), _lib.stream_ptr(x.device))
    return y


class DeformConv(nn.Module):
    """ops/dcn/deform_conv.py:190-236."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=False):
        super().__init__()
        assert not bias
        assert in_channels % groups == 0, "in_channels {} cannot be divisible by groups {}".format(in_channels, groups)
        assert out_channels % groups == 0, "out_channels {} cannot be divisible by groups {}".format(
            out_channels, groups)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = _pair(stride), _pair(padding), _pair(dilation)
        self.groups, self.deformable_groups = groups, deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight.data.uniform_(-bound, bound)

    def forward(self, x, offset):
        return deform_conv(x, offset, self.weight, self.stride, self.padding, self.dilation, self.groups,
                           self.deformable_groups)


class DeformConvPack(DeformConv):
    """ops/dcn/deform_conv.py:239-261: owns a zero-initialised offset conv."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset = nn.Conv2d(
            self.in_channels, self.deformable_groups * 2 * self.kernel_size[0] * self.kernel_size[1],
            kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        self.init_offset()

    def init_offset(self):
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        return deform_conv(x, self.conv_offset(x), self.weight, self.stride, self.padding, self.dilation,
                           self.groups, self.deformable_groups)


class ModulatedDeformConv(nn.Module):
    """ops/dcn/deform_conv.py:264-308 (scalar stride/padding/dilation)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deformable_groups = groups, deformable_groups
        self.with_bias = bias
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight.data.uniform_(-bound, bound)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)


class ModulatedDeformConvPack(ModulatedDeformConv):
    """ops/dcn/deform_conv.py:311-337: owns a zero-initialised offset+mask conv.

    Unlike the reference, `init_offset` is only invoked when this class is the
    concrete type: the reference calls it unconditionally (deform_conv.py:324),
    which makes subclasses that override it (MVDualAttAlignment,
    arch/SIDECVSR_our.py:3293-3301) impossible to construct.
    """

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_offset_mask = nn.Conv2d(
            self.in_channels, self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1],
            kernel_size=self.kernel_size, stride=_pair(self.stride), padding=_pair(self.padding), bias=True)
        if type(self).init_offset is ModulatedDeformConvPack.init_offset:
            self.init_offset()
        else:
            ModulatedDeformConvPack.init_offset(self)

    def init_offset(self):
        nn.init.zeros_(self.conv_offset_mask.weight)
        nn.init.zeros_(self.conv_offset_mask.bias)

    def forward(self, x):
        out = self.conv_offset_mask(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        return modulated_deform_conv(x, torch.cat((o1, o2), dim=1), torch.sigmoid(mask), self.weight, self.bias,
                                     self.stride, self.padding, self.dilation, self.groups, self.deformable_groups)
