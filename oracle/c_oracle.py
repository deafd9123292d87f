"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/dcn_ref.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this.  See oracle/dcn_ref.c for the reference lines restated.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_dcn.so")
_lib = None


def build(force=False):
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "dcn_ref.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_dcn_forward.restype = ctypes.c_int
        _lib.oracle_flow_warp.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1]))


def dcn_forward(x, offset, mask, weight, bias, stride=1, padding=0, dilation=1,
                groups=1, deformable_groups=1, return_index=False):
    """DCNv2 (mask given) / DCNv1 (mask None) forward, fp32 numpy NCHW."""
    x = np.ascontiguousarray(x, np.float32)
    offset = np.ascontiguousarray(offset, np.float32)
    weight = np.ascontiguousarray(weight, np.float32)
    mask = None if mask is None else np.ascontiguousarray(mask, np.float32)
    bias = None if bias is None else np.ascontiguousarray(bias, np.float32)
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    y = np.empty((B, Co, Ho, Wo), np.float32)
    idx = np.empty((B, deformable_groups * kh * kw, Ho, Wo, 2), np.int32) if return_index else None
    rc = lib().oracle_dcn_forward(
        _fp(x), _fp(offset), _fp(mask), _fp(weight), _fp(bias), _fp(y),
        B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, groups, deformable_groups, _fp(idx))
    if rc != 0:
        raise ValueError("oracle_dcn_forward: shape error %d" % rc)
    return (y, idx) if return_index else y


def flow_warp(x, flow, formula=0, return_index=False):
    """x [B,C,H,W], flow [B,H,W,2] (x,y) -> warped [B,C,H,W] (bilinear, zeros, align_corners=True)."""
    x = np.ascontiguousarray(x, np.float32)
    flow = np.ascontiguousarray(flow, np.float32)
    B, C, H, W = x.shape
    assert flow.shape == (B, H, W, 2)
    y = np.empty_like(x)
    idx = np.empty((B, H, W, 2), np.int32) if return_index else None
    lib().oracle_flow_warp(_fp(x), _fp(flow), _fp(y), B, C, H, W, int(formula), _fp(idx))
    return (y, idx) if return_index else y
