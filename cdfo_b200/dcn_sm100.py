"""Host side of the tcgen05 DCN kernel (cdfo_dcn_sm100_fwd): operand packing, weight cache, launch."""
import ctypes

import torch

from . import _lib

_wcache = _lib.TensorCache()

# When a list, every dcn_sm100() launch appends (start_event, end_event, pixels) recorded on the launching
# stream: bench.py reads the kernel's live duration from these for its roofline line.
event_log = None


def supported(x, weight, stride, padding, dilation, groups, deformable_groups, mask):
    """The model's hot shape: 64 -> 64 channels, 3x3, stride = padding = dilation = 1, one weight group."""
    return (mask is not None and x.is_cuda and x.dim() == 4 and x.size(1) == 64 and tuple(weight.shape) == (64, 64, 3, 3)
            and tuple(stride) == (1, 1) and tuple(padding) == (1, 1) and tuple(dilation) == (1, 1) and groups == 1
            and deformable_groups in (1, 2, 4, 8, 16) and x.dtype in (torch.float32, torch.bfloat16))


@torch.no_grad()
def pack_weight(weight: torch.Tensor) -> torch.Tensor:
    """[64,64,3,3] -> bf16 B operand [9, 8, 64, 8]; cached per parameter tensor / version counter."""
    return _wcache.get(weight, lambda: _pack_weight(weight))


def _pack_weight(weight):
    w = weight.detach().contiguous().float()
    out = torch.empty((9, 8, 64, 8), dtype=torch.bfloat16, device=w.device)
    _lib.call("cdfo_dcn_sm100_pack_weight", _lib.ptr(w), _lib.ptr(out), _lib.stream_ptr(w.device))
    return out


@torch.no_grad()
def pack_q4p(x: torch.Tensor) -> torch.Tensor:
    """NCHW fp32 -> [B, C/4, H+3, W+3, 4] bf16: quad-planar, zero border of 1 pixel before and 2 after."""
    B, C, H, W = x.shape
    x = x.contiguous().float()
    out = torch.empty((B, C // 4, H + 3, W + 3, 4), dtype=torch.bfloat16, device=x.device)
    _lib.call("cdfo_pack_q4p", _lib.ptr(x), _lib.ptr(out), B, C, H, W, _lib.stream_ptr(x.device))
    return out


@torch.no_grad()
def dcn_sm100(x_q4p, offset, mask, wpk, bias=None, mv=None, out_c8=False, num_ctas=0, fused_fields=None):
    """x_q4p [B,16,H+3,W+3,4] bf16; offset [B,dg*18,H,W], mask [B,dg*9,H,W] fp32 or fp16; mv [B,2,H,W] fp32 or None.
    Returns [B,64,H,W] fp32 (out_c8=False) or [B,8,H,W,8] bf16."""
    xB, _, Hp, Wp, _ = x_q4p.shape
    H, W = Hp - 3, Wp - 3
    if fused_fields is not None:
        # packed fields [B, 9, dg/gp, H, W, gp, 4] fp16 = (dy, dx, mask, 0): the fused head's output
        B, dg = _check_fields(fused_fields, H, W)
        return _launch(x_q4p, fused_fields, None, wpk, bias, mv, out_c8, num_ctas, B, H, W, dg, xB, 0, 0, FIELDS_F16X4)
    B = offset.size(0)
    dg = offset.size(1) // 18
    if offset.dtype != mask.dtype or offset.dtype not in (torch.float32, torch.float16):
        raise _lib.CdfoError("dcn_sm100: offset/mask must both be fp32 or fp16")
    if tuple(offset.shape) != (B, dg * 18, H, W) or tuple(mask.shape) != (B, dg * 9, H, W):
        raise _lib.CdfoError("dcn_sm100: offset/mask shape mismatch")
    offset, mask = offset.contiguous(), mask.contiguous()
    return _launch(x_q4p, offset, mask, wpk, bias, mv, out_c8, num_ctas, B, H, W, dg, xB, 0, 0)


FIELDS_F16X4 = 16  # CDFO_FIELDS_F16X4


def _launch(x_q4p, offset, mask, wpk, bias, mv, out_c8, num_ctas, B, H, W, dg, xB, off_bs, msk_bs, off_code=None):
    if B % xB:
        raise _lib.CdfoError("dcn_sm100: batch %d is not a multiple of the x batch %d" % (B, xB))
    if mv is not None:
        mv = mv.contiguous().float()
    if bias is not None:
        bias = bias.detach().contiguous().float()
    if out_c8:
        y = torch.empty((B, 8, H, W, 8), dtype=torch.bfloat16, device=x_q4p.device)
    else:
        y = torch.empty((B, 64, H, W), dtype=torch.float32, device=x_q4p.device)
    ev = None
    if event_log is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _lib.call("cdfo_dcn_sm100_fwd", 
        _lib.ptr(x_q4p), _lib.ptr(offset), _lib.ptr(mask), _lib.ptr(mv), _lib.ptr(wpk), _lib.ptr(bias), _lib.ptr(y),
        B, H, W, dg, _lib.dtype_code(offset) if off_code is None else off_code, 1 if out_c8 else 0, int(num_ctas), int(xB),
        ctypes.c_longlong(off_bs), ctypes.c_longlong(msk_bs), _lib.stream_ptr(x_q4p.device))
    if ev is not None:
        ev[1].record()
        event_log.append((ev[0], ev[1], B * H * W, 8 if off_code == FIELDS_F16X4 else offset.element_size()))
    return y


# ------------------------------------------------------------------------------------------ texture-unit gather (v3)
_wcache16 = _lib.TensorCache()


@torch.no_grad()
def pack_weight_f16(weight: torch.Tensor) -> torch.Tensor:
    """[64,64,3,3] -> fp16 B operand [9, 8, 64, 8] of cdfo_dcn_tex_sm100_fwd; cached per parameter / version."""
    return _wcache16.get(weight, lambda: _pack_weight_f16(weight))


def _pack_weight_f16(weight):
    w = weight.detach().contiguous().float()
    out = torch.empty((9, 8, 64, 8), dtype=torch.float16, device=w.device)
    _lib.call("cdfo_dcn_tex_sm100_pack_weight", _lib.ptr(w), _lib.ptr(out), _lib.stream_ptr(w.device))
    return out


@torch.no_grad()
def pack_q4t(x: torch.Tensor) -> torch.Tensor:
    """NCHW fp32 -> [B, C/4, H+3, Wpt, 4] fp16 texels (zero border 1 before / 2 after, pitch padded to 32 bytes)."""
    B, C, H, W = x.shape
    x = x.contiguous().float()
    wpt = _lib.lib().cdfo_q4t_pitch(W)
    out = torch.empty((B, C // 4, H + 3, wpt, 4), dtype=torch.float16, device=x.device)
    _lib.call("cdfo_pack_q4t", _lib.ptr(x), _lib.ptr(out), B, C, H, W, _lib.stream_ptr(x.device))
    return out


@torch.no_grad()
def dcn_tex(x_q4t, fields, wpk16, bias=None, mv=None, out_c8=False, num_ctas=0):
    """x_q4t [xB,16,H+3,Wpt,4] fp16; fields [B, 9, H, W, dg, 4] fp16 (dy, dx, mask, 0); mv [B,2,H,W] fp32 or None.
    Returns [B,64,H,W] fp32 (out_c8=False) or [B,8,H,W,8] bf16."""
    xB, _, Hp, _, _ = x_q4t.shape
    H, W = fields.size(3), fields.size(4)
    B, dg = _check_fields(fields, H, W)
    if H != Hp - 3 or x_q4t.size(3) != _lib.lib().cdfo_q4t_pitch(W) or x_q4t.dtype != torch.float16 or not x_q4t.is_contiguous():
        raise _lib.CdfoError("dcn_tex: x_q4t does not match the %dx%d fields (use pack_q4t)" % (H, W))
    if B % xB:
        raise _lib.CdfoError("dcn_tex: batch %d is not a multiple of the x batch %d" % (B, xB))
    if mv is not None:
        mv = mv.contiguous().float()
    if bias is not None:
        bias = bias.detach().contiguous().float()
    if out_c8:
        y = torch.empty((B, 8, H, W, 8), dtype=torch.bfloat16, device=fields.device)
    else:
        y = torch.empty((B, 64, H, W), dtype=torch.float32, device=fields.device)
    ev = None
    if event_log is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _lib.call("cdfo_dcn_tex_sm100_fwd", _lib.ptr(x_q4t), _lib.ptr(fields), _lib.ptr(mv), _lib.ptr(wpk16), _lib.ptr(bias),
              _lib.ptr(y), B, H, W, int(dg), 1 if out_c8 else 0, int(num_ctas), int(xB), ctypes.c_longlong(0),
              _lib.stream_ptr(fields.device))
    if ev is not None:
        ev[1].record()
        event_log.append((ev[0], ev[1], B * H * W, 8))
    return y


@torch.no_grad()
def dcn_tex_stacked(x_q4t, fields, wpk16, bias, mv, stack, group_chunk):
    """dcn_tex writing into `stack` [n_seq, chunks, H, W, 8] bf16: sample s = g * n_seq + b of the group-major batch lands in
    chunks [group_chunk[g], +8) of stack[b] (cdfo_dcn_tex_sm100_stacked_fwd)."""
    xB = x_q4t.size(0)
    H, W = fields.size(3), fields.size(4)
    B, dg = _check_fields(fields, H, W)
    n_seq, chunks = stack.size(0), stack.size(1)
    if fields.dtype != torch.float16 or not fields.is_contiguous() or stack.dtype != torch.bfloat16 or not stack.is_contiguous() or \
            tuple(stack.shape[2:]) != (H, W, 8) or B % n_seq or len(group_chunk) != B // n_seq:
        raise _lib.CdfoError("dcn_tex_stacked: fields / stack mismatch")
    grp = (ctypes.c_int * len(group_chunk))(*[int(g) for g in group_chunk])
    ev = None
    if event_log is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _lib.call("cdfo_dcn_tex_sm100_stacked_fwd", _lib.ptr(x_q4t), _lib.ptr(fields), _lib.ptr(None if mv is None else mv.contiguous().float()),
              _lib.ptr(wpk16), _lib.ptr(None if bias is None else bias.detach().contiguous().float()), _lib.ptr(stack),
              int(n_seq), len(group_chunk), int(chunks), grp, H, W, int(dg), int(xB), _lib.stream_ptr(fields.device))
    if ev is not None:
        ev[1].record()
        event_log.append((ev[0], ev[1], B * H * W, 8))
    return None


def _check_fields(fields, H, W):
    """fields [B, 9, dg/gp, H, W, gp, 4] fp16 with gp = 2 when dg = 16, else 1 -> (B, dg)."""
    if fields.dtype != torch.float16 or not fields.is_contiguous() or fields.dim() != 7 or fields.size(1) != 9 or \
            tuple(fields.shape[3:5]) != (H, W) or fields.size(6) != 4:
        raise _lib.CdfoError("fields must be a contiguous fp16 [B, 9, dg/gp, H, W, gp, 4] tensor")
    dg = fields.size(2) * fields.size(5)
    if fields.size(5) != (2 if dg == 16 else 1):
        raise _lib.CdfoError("fields: group pairing must be 2 for 16 deformable groups and 1 otherwise")
    return fields.size(0), dg


def fields_shape(B, dg, H, W):
    gp = 2 if dg == 16 else 1
    return (B, 9, dg // gp, H, W, gp, 4)


def pack_fields(offset, mask, dg):
    """Reference-layout offset [B, dg*18, H, W] (channel 2*(g*9+tap) + {0: dy, 1: dx}) and mask [B, dg*9, H, W] ->
    packed fields [B, 9, dg/gp, H, W, gp, 4] fp16 = (dy, dx, mask, 0), group g = pair * gp + e."""
    B, _, H, W = offset.shape
    gp = 2 if dg == 16 else 1
    o = offset.reshape(B, dg // gp, gp, 9, 2, H, W).permute(0, 3, 1, 5, 6, 2, 4)          # [B, 9, dg/gp, H, W, gp, 2]
    m = mask.reshape(B, dg // gp, gp, 9, H, W).permute(0, 3, 1, 4, 5, 2).unsqueeze(-1)     # [B, 9, dg/gp, H, W, gp, 1]
    return torch.cat([o, m, torch.zeros_like(m)], dim=-1).half().contiguous()


def unpack_fields(fields):
    """Packed fields -> (offset [B, dg*18, H, W], mask [B, dg*9, H, W]) fp32 in the reference's channel order."""
    B, _, npair, H, W, gp, _ = fields.shape
    dg = npair * gp
    f = fields.float()
    off = f[..., :2].permute(0, 2, 5, 1, 6, 3, 4).reshape(B, dg * 18, H, W).contiguous()   # [B, pair, gp, 9, 2, H, W]
    msk = f[..., 2].permute(0, 2, 5, 1, 3, 4).reshape(B, dg * 9, H, W).contiguous()
    return off, msk


# ------------------------------------------------------------------------------------------ fused head + DCN (csrc/mv_dcn_fused_sm100.cu)
def fused_head_permutation(dg, device):
    """Output-channel order of conv_offset[-1] the fused kernel expects: n' = T*144 + qs*36 + tl*12 + gi*3 + c  <-  reference channel
    (2k, 2k+1, dg*18 + k)[c] with k = g*9 + tap, tap = 3T + tl, g = 4 qs + gi (the chunk / cat of arch:3341-3345)."""
    if dg != 16:
        raise _lib.CdfoError("fused head + DCN: 16 deformable groups only")
    n = torch.arange(432, device=device)
    T, r = n // 144, n % 144
    qs, r = r // 36, r % 36
    tl, r = r // 12, r % 12
    gi, c = r // 3, r % 3
    k = (4 * qs + gi) * 9 + 3 * T + tl
    return torch.where(c == 2, dg * 18 + k, 2 * k + c)


@torch.no_grad()
def mv_head_dcn_fused(z8, head_wpk, head_bias, magnitude, x_q4t, mv, wpk16, bias=None, out_c8=False, stack=None, group_chunk=None,
                      fields_out=None, num_ctas=0):
    """z8 [2B, 8, H, W, 8] bf16 hidden maps (sample b and b + B); x_q4t [xB, 16, H+3, Wpt, 4] fp16; mv [B, 2, H, W] fp32 or None.
    Returns [B, 64, H, W] fp32, [B, 8, H, W, 8] bf16 (out_c8), or None when writing into `stack` ([n_seq, chunks, H, W, 8] bf16,
    group-major batch like dcn_tex_stacked).  fields_out: optional fp16 [B, 9, 8, H, W, 2, 4] debug tap of the fields."""
    B2, c8, H, W, e = z8.shape
    if z8.dtype != torch.bfloat16 or not z8.is_contiguous() or c8 != 8 or e != 8 or B2 % 2:
        raise _lib.CdfoError("mv_head_dcn_fused: z8 must be a contiguous bf16 [2B, 8, H, W, 8] tensor")
    B, xB = B2 // 2, x_q4t.size(0)
    if x_q4t.size(2) != H + 3 or x_q4t.size(3) != _lib.lib().cdfo_q4t_pitch(W) or x_q4t.dtype != torch.float16 or not x_q4t.is_contiguous():
        raise _lib.CdfoError("mv_head_dcn_fused: x_q4t does not match the %dx%d hidden maps (use pack_q4t)" % (H, W))
    if B % xB:
        raise _lib.CdfoError("mv_head_dcn_fused: batch %d is not a multiple of the x batch %d" % (B, xB))
    if head_bias.dtype != torch.float32 or head_bias.numel() != 432 or not head_bias.is_contiguous():
        raise _lib.CdfoError("mv_head_dcn_fused: head bias must be 432 contiguous fp32 values")
    mv = None if mv is None else mv.contiguous().float()
    bias = None if bias is None else bias.detach().contiguous().float()
    dev = z8.device
    ev = None
    if event_log is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    if stack is not None:
        n_seq, chunks = stack.size(0), stack.size(1)
        if stack.dtype != torch.bfloat16 or not stack.is_contiguous() or tuple(stack.shape[2:]) != (H, W, 8) or B % n_seq or \
                len(group_chunk) != B // n_seq:
            raise _lib.CdfoError("mv_head_dcn_fused: stack mismatch")
        grp = (ctypes.c_int * len(group_chunk))(*[int(g) for g in group_chunk])
        _lib.call("cdfo_mv_head_dcn_fused_sm100_stacked_fwd", _lib.ptr(z8), _lib.ptr(head_wpk), _lib.ptr(head_bias),
                  ctypes.c_float(float(magnitude)), _lib.ptr(x_q4t), _lib.ptr(mv), _lib.ptr(wpk16), _lib.ptr(bias), _lib.ptr(stack),
                  int(n_seq), len(group_chunk), int(chunks), grp, H, W, int(xB), _lib.stream_ptr(dev))
        y = None
    else:
        if out_c8:
            y = torch.empty((B, 8, H, W, 8), dtype=torch.bfloat16, device=dev)
        else:
            y = torch.empty((B, 64, H, W), dtype=torch.float32, device=dev)
        if fields_out is not None and (fields_out.dtype != torch.float16 or not fields_out.is_contiguous() or
                                       tuple(fields_out.shape) != fields_shape(B, 16, H, W)):
            raise _lib.CdfoError("mv_head_dcn_fused: fields_out must be a contiguous fp16 %s tensor" % (fields_shape(B, 16, H, W),))
        _lib.call("cdfo_mv_head_dcn_fused_sm100_fwd", _lib.ptr(z8), _lib.ptr(head_wpk), _lib.ptr(head_bias),
                  ctypes.c_float(float(magnitude)), _lib.ptr(x_q4t), _lib.ptr(mv), _lib.ptr(wpk16), _lib.ptr(bias), _lib.ptr(y),
                  _lib.ptr(fields_out), B, H, W, 1 if out_c8 else 0, int(xB), int(num_ctas), _lib.stream_ptr(dev))
    if ev is not None:
        ev[1].record()
        event_log.append((ev[0], ev[1], B * H * W, "fused"))
    return y
