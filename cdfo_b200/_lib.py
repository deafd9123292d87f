"""ctypes binding of libcdfo_b200.so (the C ABI declared in include/cdfo_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is
missing, or a call returns a non-zero status, this raises.
"""
import ctypes
import os
import weakref

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcdfo_b200.so")

OK = 0
F32, F16, BF16 = 0, 1, 2
_DTYPES = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}


class CdfoError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(SO_PATH):
            raise CdfoError(
                "cdfo_b200: %s not found -- build it with `python cdfo_b200/csrc/build.py` "
                "(there is no CPU or PyTorch fallback for this path)" % SO_PATH)
        _lib = ctypes.CDLL(SO_PATH)
        _lib.cdfo_last_error.restype = ctypes.c_char_p
        _lib.cdfo_conv3x3_sm100_weight_bytes.restype = ctypes.c_size_t
        _lib.cdfo_conv_sm100_weight_bytes.restype = ctypes.c_size_t
        _lib.cdfo_lra_workspace_bytes.restype = ctypes.c_size_t
        _lib.cdfo_q4t_bytes.restype = ctypes.c_size_t
        _lib.cdfo_mdta_workspace_bytes.restype = ctypes.c_size_t
        _lib.cdfo_psnr_ssim_workspace_bytes.restype = ctypes.c_size_t
        _lib.cdfo_conv3x3_pair_sm100_weight_bytes.restype = ctypes.c_size_t
        _lib.cdfo_conv4x4s2_pair_sm100_weight_bytes.restype = ctypes.c_size_t
        _lib.cdfo_lra_mask_logits_workspace_bytes.restype = ctypes.c_size_t
    return _lib


# number of CUDA kernels each C-ABI entry launches (for bench.py's gpu_launches claim)
_LAUNCHES = {"cdfo_mv_end_fix": 3, "cdfo_lra_fwd": 5, "cdfo_lra_c8_fwd": 5, "cdfo_mdta_fwd": 3, "cdfo_psnr_ssim_u8": 3, "cdfo_lra_mask_logits_fwd": 2,
             "cdfo_spatial_gate_c8_fwd": 2}
launch_count = 0


def call(name, *args):
    """Invoke a C-ABI entry point, count its kernel launches, raise on a non-zero status."""
    global launch_count
    rc = getattr(lib(), name)(*args)
    if rc != OK:
        check(rc, name)
    launch_count += _LAUNCHES.get(name, 1)
    return rc


def check(rc, what=""):
    if rc != OK:
        msg = lib().cdfo_last_error().decode("utf-8", "replace")
        # same exception classes as the reference raises at this boundary
        # (TORCH_CHECK / AT_ERROR -> RuntimeError, deform_conv_cuda.cpp:493-511)
        raise CdfoError("%s failed (status %d): %s" % (what or "cdfo call", rc, msg))


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise CdfoError("cdfo_b200: unsupported dtype %s" % t.dtype)


def ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            # the reference raises NotImplementedError for CPU tensors (deform_conv.py:46-47,136-137)
            raise NotImplementedError("cdfo_b200 ops are CUDA-only (got a %s tensor)" % t.device)


class TensorCache:
    """Derived tensors (packed / permuted weights) keyed by the SOURCE tensor object: an entry is valid only while that very object is
    alive (weak reference identity, never id() alone -- CPython reuses freed addresses) with the same storage address, device and
    in-place version; entries of collected tensors are dropped by the weak reference's callback, so nothing outlives its model."""

    def __init__(self):
        self._d = {}

    def get(self, t, build, kind=None):
        key = (id(t), kind)
        sig = (t.data_ptr(), str(t.device), t._version, t.dtype)
        hit = self._d.get(key)
        if hit is not None and hit[0]() is t and hit[1] == sig:
            return hit[2]
        val = build()
        d = self._d
        self._d[key] = (weakref.ref(t, lambda _r, k=key: d.pop(k, None)), sig, val)
        return val

    def __len__(self):
        return len(self._d)
