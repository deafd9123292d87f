"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: CPU-only box)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(cdfo_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_something():
    syms = declared_symbols()
    assert "cdfo_dcn_fwd" in syms and "cdfo_last_error" in syms and len(syms) >= 8


def test_library_exports_every_declared_symbol(cdfo_so):
    lib = ctypes.CDLL(cdfo_so)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in include/ but not exported: %s" % missing


def test_error_plumbing_without_gpu(cdfo_so):
    lib = ctypes.CDLL(cdfo_so)
    lib.cdfo_last_error.restype = ctypes.c_char_p
    assert lib.cdfo_version() >= 1000
    # argument validation happens before any CUDA call, so it is checkable on the CPU box
    rc = lib.cdfo_dcn_fwd(None, None, None, None, None, None, 1, 4, 8, 8, 4, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 0, None)
    assert rc == -2 and b"non-NULL" in lib.cdfo_last_error()
    buf = ctypes.create_string_buffer(16)
    rc = lib.cdfo_dcn_fwd(buf, buf, None, buf, None, buf, 1, 6, 8, 8, 4, 3, 3, 1, 1, 1, 1, 1, 1, 1, 4, 0, None)
    assert rc == -1 and b"deformable group" in lib.cdfo_last_error()


def test_product_has_no_oracle_import():
    """The shipped package must never route through oracle/ (or any CPU fallback)."""
    for path in glob.glob(os.path.join(ROOT, "cdfo_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path


def test_missing_library_fails_loudly(monkeypatch):
    import cdfo_b200._lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "SO_PATH", "/nonexistent/libcdfo_b200.so")
    try:
        L.lib()
    except L.CdfoError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("expected CdfoError")


def test_ctypes_prototypes_come_from_the_header(cdfo_so):
    """cdfo_b200._lib binds restype / argtypes of EVERY declared entry point from include/cdfo_b200.h: no default-int conversion of a
    pointer or float argument, a float where the ABI says int raises."""
    import cdfo_b200._lib as L
    protos = L.prototypes()
    assert sorted(protos) == declared_symbols()
    lib = L.lib()
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        assert fn.restype is res and list(fn.argtypes) == args, name
    assert protos["cdfo_last_error"][0] is ctypes.c_char_p and protos["cdfo_mdta_workspace_bytes"][0] is ctypes.c_size_t
    assert protos["cdfo_lra_c8_fwd"][1][6] is ctypes.c_float and protos["cdfo_dcn_fwd"][1][0] is ctypes.c_void_p
    try:
        lib.cdfo_mdta_workspace_bytes(1.5, 8, 8, 8)
    except ctypes.ArgumentError:
        pass
    else:
        raise AssertionError("a float passed for an int parameter must raise")
