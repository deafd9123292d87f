"""ctypes binding of libcdfo_b200.so (the C ABI declared in include/cdfo_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is
missing, or a call returns a non-zero status, this raises.
"""
import ctypes
import os
import re
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
HEADER = os.path.join(os.path.dirname(_HERE), "include", "cdfo_b200.h")


def _ctype(decl):
    """C parameter / return type of include/cdfo_b200.h -> ctypes type."""
    d = decl.strip()
    if "*" in d:
        return ctypes.c_void_p
    base = re.sub(r"\b(const|unsigned|signed)\b", "", d).split()
    t = base[0] if base else "int"
    return {"float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t, "int": ctypes.c_int, "int64_t": ctypes.c_int64,
            "uint8_t": ctypes.c_uint8, "long": ctypes.c_long}.get(t, ctypes.c_int)


def prototypes(header=HEADER):
    """{name: (restype, [argtypes])} of every function include/cdfo_b200.h declares: the ctypes calls are bound to the header's own
    prototypes, so a float passed where the ABI says int (or a 64-bit value where it says int) raises instead of being truncated."""
    src = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for ret, name, params in re.findall(r"\b(const\s+char\s*\*|size_t|int|void)\s*(cdfo_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src, flags=re.S):
        params = " ".join(params.split())
        args = [] if params in ("", "void") else [_ctype(a) for a in params.split(",")]
        res = ctypes.c_char_p if "char" in ret else (ctypes.c_size_t if ret == "size_t" else (None if ret == "void" else ctypes.c_int))
        out[name] = (res, args)
    return out


def _bind_prototypes(handle):
    if not os.path.isfile(HEADER):
        raise CdfoError("cdfo_b200: %s not found (the ctypes prototypes are taken from it)" % HEADER)
    for name, (res, args) in prototypes().items():
        fn = getattr(handle, name, None)
        if fn is None:
            raise CdfoError("cdfo_b200: %s declares %s but %s does not export it (stale build?)" % (HEADER, name, SO_PATH))
        fn.restype, fn.argtypes = res, args


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(SO_PATH):
            raise CdfoError(
                "cdfo_b200: %s not found -- build it with `python cdfo_b200/csrc/build.py` "
                "(there is no CPU or PyTorch fallback for this path)" % SO_PATH)
        _lib = ctypes.CDLL(SO_PATH)
        _bind_prototypes(_lib)
    return _lib


# number of CUDA kernels each C-ABI entry launches (for bench.py's gpu_launches claim)
_LAUNCHES = {"cdfo_mv_end_fix": 3, "cdfo_lra_fwd": 5, "cdfo_lra_c8_fwd": 5, "cdfo_mdta_fwd": 3, "cdfo_mdta_c8_fwd": 3, "cdfo_psnr_ssim_u8": 3,
             "cdfo_lra_mask_logits_fwd": 2,
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
