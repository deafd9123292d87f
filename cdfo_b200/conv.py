"""Host side of the tcgen05 3x3 convolution (cdfo_conv3x3_sm100_fwd) and of the c8 layout adapters."""
import weakref

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
_wcache = {}


@torch.no_grad()
def to_c8(x: torch.Tensor) -> torch.Tensor:
    """NCHW fp32 -> [B, C/8, H, W, 8] bf16."""
    B, C, H, W = x.shape
    x = x.contiguous().float()
    out = torch.empty((B, C // 8, H, W, 8), dtype=torch.bfloat16, device=x.device)
    _lib.call("cdfo_pack_c8", _lib.ptr(x), _lib.ptr(out), B, C, H, W, _lib.stream_ptr(x.device))
    return out


@torch.no_grad()
def from_c8(x8: torch.Tensor) -> torch.Tensor:
    """[B, C/8, H, W, 8] bf16 -> NCHW fp32."""
    B, C8, H, W, _ = x8.shape
    out = torch.empty((B, C8 * 8, H, W), dtype=torch.float32, device=x8.device)
    _lib.call("cdfo_unpack_c8", _lib.ptr(x8), _lib.ptr(out), B, C8 * 8, H, W, _lib.stream_ptr(x8.device))
    return out


@torch.no_grad()
def pack_weight(weight: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> packed bf16 B operand (cached per parameter / version)."""
    key = id(weight)
    hit = _wcache.get(key)
    if hit is not None and hit[0]() is weight and hit[1] == weight._version:
        return hit[2]
    Cout, Cin = weight.shape[:2]
    nbytes = _lib.lib().cdfo_conv3x3_sm100_weight_bytes(Cout, Cin)
    if nbytes == 0 or Cin % 64 or Cout % 16:
        raise _lib.CdfoError("conv3x3_sm100: unsupported channels %d -> %d" % (Cin, Cout))
    w = weight.detach().contiguous().float()
    out = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
    _lib.call("cdfo_conv3x3_sm100_pack_weight", _lib.ptr(w), _lib.ptr(out), Cout, Cin, _lib.stream_ptr(w.device))
    _wcache[key] = (weakref.ref(weight), weight._version, out)
    return out


@torch.no_grad()
def conv3x3(x8, weight, bias=None, act=ACT_NONE, resid8=None, out_nchw=False):
    """x8 [B, Cin/8, H, W, 8] bf16 -> [B, Cout/8, H, W, 8] bf16 (or [B, Cout, H, W] fp32 when out_nchw)."""
    B, C8, H, W, _ = x8.shape
    Cout, Cin = weight.shape[:2]
    if C8 * 8 != Cin or tuple(weight.shape[2:]) != (3, 3):
        raise _lib.CdfoError("conv3x3: weight %s does not match input with %d channels" % (tuple(weight.shape), C8 * 8))
    if x8.dtype != torch.bfloat16 or not x8.is_contiguous():
        raise _lib.CdfoError("conv3x3: input must be a contiguous bf16 c8 tensor")
    wpk = pack_weight(weight)
    b = None if bias is None else bias.detach().contiguous().float()
    if out_nchw:
        y = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x8.device)
    else:
        y = torch.empty((B, Cout // 8, H, W, 8), dtype=torch.bfloat16, device=x8.device)
    if resid8 is not None and (resid8.shape != (B, Cout // 8, H, W, 8) or resid8.dtype != torch.bfloat16):
        raise _lib.CdfoError("conv3x3: residual must be a bf16 c8 tensor of the output shape")
    _lib.call("cdfo_conv3x3_sm100_fwd", _lib.ptr(x8), _lib.ptr(wpk), _lib.ptr(b), _lib.ptr(resid8), _lib.ptr(y),
              B, Cin, Cout, H, W, int(act), 0 if out_nchw else 1, _lib.stream_ptr(x8.device))
    return y
