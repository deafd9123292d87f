"""Host side of the frame I/O conversions and the on-GPU PSNR / SSIM (csrc/metrics.cu).

Mirrors what the reference's eval loop does on the CPU: generate_input (test_LD_37.py:19-29), the uint8 conversion before
cv2.imwrite (:172-180) and calculate_psnr / calculate_ssim as cal_psnr_ssim calls them (metric/psnr_ssim.py:278-399,446-484).
"""
import torch

from . import _lib

_KINDS = {torch.uint8: 0, torch.int8: 1, torch.int16: 2, torch.int32: 3}


@torch.no_grad()
def planes_to_unit(planes: torch.Tensor, rows_out=None, out=None) -> torch.Tensor:
    """Integer planes [..., H, W] (uint8 / int8 / int16 / int32) -> float32 k / 255 [..., rows_out, W], extra rows zero."""
    _lib.require_cuda(planes)
    if planes.dtype not in _KINDS:
        raise _lib.CdfoError("planes_to_unit: unsupported dtype %s" % planes.dtype)
    planes = planes.contiguous()
    H, W = planes.shape[-2:]
    rows_out = H if rows_out is None else int(rows_out)
    n = planes.numel() // (H * W)
    shape = tuple(planes.shape[:-2]) + (rows_out, W)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=planes.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise _lib.CdfoError("planes_to_unit: out must be a contiguous float32 %s tensor" % (shape,))
    _lib.call("cdfo_planes_to_unit_f32", _lib.ptr(planes), _KINDS[planes.dtype], _lib.ptr(out), n, H, W, rows_out,
              _lib.stream_ptr(planes.device))
    return out


@torch.no_grad()
def sr_to_u8(sr: torch.Tensor, rows_out=None, out=None) -> torch.Tensor:
    """SR float32 [..., H, W] -> uint8 [..., rows_out, W]: drop the padded rows, clamp(0, 1) * 255, truncate."""
    _lib.require_cuda(sr)
    sr = sr.contiguous().float()
    H, W = sr.shape[-2:]
    rows_out = H if rows_out is None else int(rows_out)
    shape = tuple(sr.shape[:-2]) + (rows_out, W)
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=sr.device)
    elif tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous():
        raise _lib.CdfoError("sr_to_u8: out must be a contiguous uint8 %s tensor" % (shape,))
    _lib.call("cdfo_sr_to_u8", _lib.ptr(sr), _lib.ptr(out), sr.numel() // (H * W), H, W, rows_out, _lib.stream_ptr(sr.device))
    return out


_ws = {}


@torch.no_grad()
def psnr_ssim(res: torch.Tensor, gt: torch.Tensor, crop_border=4, accum=None):
    """res, gt uint8 [B, H, W] (or [B, 1, H, W]) -> float64 [B, 2] = (PSNR dB, SSIM) per frame on the Y channel.
    accum (float64 [B, 3], optional) += (psnr, ssim, 1): the sums cal_psnr_ssim divides by the frame count."""
    _lib.require_cuda(res, gt)
    if res.dtype != torch.uint8 or gt.dtype != torch.uint8 or res.shape != gt.shape:
        raise _lib.CdfoError("psnr_ssim: two uint8 tensors of the same shape expected")   # metric/psnr_ssim.py:296-297
    H, W = res.shape[-2:]
    B = res.numel() // (H * W)
    res, gt = res.contiguous(), gt.contiguous()
    need = _lib.lib().cdfo_psnr_ssim_workspace_bytes(B, H, W, int(crop_border))
    if need == 0:
        raise _lib.CdfoError("psnr_ssim: %dx%d frame with border %d is smaller than the 11x11 SSIM window" % (H, W, crop_border))
    key = (res.device, torch.cuda.current_stream(res.device).cuda_stream)
    ws = _ws.get(key)
    if ws is None or ws.numel() < need:
        ws = _ws[key] = torch.empty(need, dtype=torch.uint8, device=res.device)
    if accum is not None and (accum.dtype != torch.float64 or tuple(accum.shape) != (B, 3) or not accum.is_contiguous()):
        raise _lib.CdfoError("psnr_ssim: accum must be a contiguous float64 [B, 3] tensor")
    out = torch.empty((B, 2), dtype=torch.float64, device=res.device)
    _lib.call("cdfo_psnr_ssim_u8", _lib.ptr(res), _lib.ptr(gt), B, H, W, int(crop_border), _lib.ptr(out), _lib.ptr(accum),
              _lib.ptr(ws), _lib.stream_ptr(res.device))
    return out
