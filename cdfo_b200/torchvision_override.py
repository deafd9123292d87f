"""CUDA-key override of `torchvision::deform_conv2d` (the second drop-in point -- ZERO changes to arch/SIDECVSR_our.py).

The DCN alignment of the reference does not call its vendored op but `torchvision.ops.deform_conv2d` (arch/SIDECVSR_our.py:3164-3165,
3260-3261, 3352, 3733-3734), whose Python wrapper (torchvision/ops/deform_conv.py:63-107) dispatches to

    torch.ops.torchvision.deform_conv2d(input, weight, offset, mask, bias, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w,
                                        n_weight_grps, n_offset_grps, use_mask) -> Tensor

`install()` registers this library's forward as that operator's CUDA kernel through torch.library, so every call site of the
unmodified reference (and anything else in the process) lands in cdfo_dcn_fwd / the tcgen05 kernel for CUDA tensors; the CPU kernel,
the Autocast and the Autograd registrations of torchvision are left alone (the autograd wrapper redispatches to this kernel for
the forward; its backward is still torchvision's, which does not need this forward's internals).  `uninstall()` drops the override.
"""
import torch

from . import _lib, config, dcn, dcn_sm100

_handle = None


def _deform_conv2d_cuda(input, weight, offset, mask, bias, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, n_weight_grps,
                        n_offset_grps, use_mask):
    """Same contract as torchvision's deform_conv2d_forward_kernel (torchvision/csrc/ops/cuda/deform_conv2d_kernel.cu): `mask` is a
    dummy [B, 0, ...]-like tensor when use_mask is false (DCNv1), `bias` always a [Co] tensor."""
    x = input.contiguous()
    B, C, H, W = x.shape
    Co, Ck, kh, kw = weight.shape
    if C != Ck * n_weight_grps or Co % n_weight_grps:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (C, Ck * n_weight_grps))
    oh, ow = dcn._out_hw((H, W), (kh, kw), (stride_h, stride_w), (pad_h, pad_w), (dil_h, dil_w))
    if tuple(offset.shape) != (B, n_offset_grps * 2 * kh * kw, oh, ow):
        raise RuntimeError("offset.shape is not valid: got %s expected %s" % (tuple(offset.shape), (B, n_offset_grps * 2 * kh * kw, oh, ow)))
    m = mask if use_mask else None
    if m is not None and tuple(m.shape) != (B, n_offset_grps * kh * kw, oh, ow):
        raise RuntimeError("mask.shape is not valid: got %s expected %s" % (tuple(m.shape), (B, n_offset_grps * kh * kw, oh, ow)))
    cast = lambda t: None if t is None else t.to(x.dtype).contiguous()  # noqa: E731
    offset, m, weight, bias = cast(offset), cast(m), cast(weight), cast(bias)
    with torch.no_grad():
        if config.tensor_core and dcn_sm100.supported(x, weight, (stride_h, stride_w), (pad_h, pad_w), (dil_h, dil_w), n_weight_grps,
                                                      n_offset_grps, m):
            return dcn._tensor_core_modulated(x, offset, m, weight, bias, n_offset_grps)
        y = x.new_empty((B, Co, oh, ow))
        _lib.call("cdfo_dcn_fwd", _lib.ptr(x), _lib.ptr(offset), _lib.ptr(m), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(y),
                  B, C, H, W, Co, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, n_weight_grps, n_offset_grps,
                  _lib.dtype_code(x), _lib.stream_ptr(x.device))
    return y


def install():
    """Register the override (idempotent).  Returns True when it is active."""
    global _handle
    if _handle is not None:
        return True
    import torchvision  # noqa: F401  (defines the torchvision::deform_conv2d schema)
    lib = torch.library.Library("torchvision", "IMPL")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # "Overriding a previously registered kernel": that is the point
        lib.impl("deform_conv2d", _deform_conv2d_cuda, "CUDA")
    _handle = lib
    return True


def uninstall():
    """Drop the override: torchvision's own CUDA kernel is active again."""
    global _handle
    if _handle is not None:
        _handle._destroy()
        _handle = None


def installed():
    return _handle is not None
