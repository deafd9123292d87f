"""TEST INFRASTRUCTURE ONLY -- import the real reference (read-only /root/reference).

Used by `oracle/make_golden.py` (run in the build container, where
/root/reference exists) to produce the committed fixtures under
tests/golden/.  Nothing under cdfo_b200/ may import this file, and nothing at
GPU run time may: /root/reference does not exist on the GPU box.

The reference does not import as shipped; the shims below are the minimum that
makes `arch/SIDECVSR_our.py` importable on a CPU-only box without touching it:

  * `timm.models.layers` (arch/SIDECVSR_our.py:8,33) and `matplotlib.pylab`
    (:34) are not installed -> empty stand-ins (DropPath/to_2tuple/
    trunc_normal_ are only used by dead classes).
  * `deform_conv_cuda` (ops/dcn/deform_conv.py:11) is a compiled module that is
    not shipped -> empty stand-in (the live model never calls it; the DCN
    alignment calls torchvision.ops.deform_conv2d, arch/SIDECVSR_our.py:3352).
  * `arch.ops.dcn` (arch/SIDECVSR_our.py:9) does not exist (ops/ lives at the
    top level) -> alias of ops.dcn.deform_conv.ModulatedDeformConv; the Pack
    class cannot be the base of MVDualAttAlignment because its __init__ calls
    the overridden init_offset() before conv_offset exists
    (ops/dcn/deform_conv.py:324 vs arch/SIDECVSR_our.py:3270-3301).
  * LLongRangAttention.__init__ calls .cuda() (arch/SIDECVSR_our.py:2161-2162)
    -> nn.Module.cuda is neutralised while the model is constructed.
  * featuremap_visual writes PNGs to a hard-coded path from inside forward
    (arch/SIDECVSR_our.py:4450,4455,4472,4475) -> no-op.
"""
import contextlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("CDFO_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "arch", "SIDECVSR_our.py"))


_A = None


def import_reference_arch():
    """Returns the reference module arch.SIDECVSR_our (cached)."""
    global _A
    if _A is not None:
        return _A
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    tl = types.ModuleType("timm.models.layers")
    tl.DropPath = type("DropPath", (nn.Identity,), {})
    tl.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    tl.trunc_normal_ = nn.init.trunc_normal_
    for name, mod in (("timm", types.ModuleType("timm")),
                      ("timm.models", types.ModuleType("timm.models")),
                      ("timm.models.layers", tl)):
        sys.modules.setdefault(name, mod)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.pylab = types.ModuleType("matplotlib.pylab")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pylab"] = mpl.pylab
    sys.modules.setdefault("deform_conv_cuda", types.ModuleType("deform_conv_cuda"))
    import ops.dcn.deform_conv as dc  # reference module, unmodified
    shim = types.ModuleType("arch.ops.dcn")
    shim.ModulatedDeformConvPack = dc.ModulatedDeformConv
    aops = types.ModuleType("arch.ops")
    aops.dcn = shim
    sys.modules["arch.ops"] = aops
    sys.modules["arch.ops.dcn"] = shim
    import arch.SIDECVSR_our as A
    A.featuremap_visual = lambda *a, **k: None
    _A = A
    return A


@contextlib.contextmanager
def cpu_construction():
    """Neutralise nn.Module.cuda while a reference module is constructed on CPU."""
    saved = nn.Module.cuda
    nn.Module.cuda = lambda self, device=None: self
    try:
        yield
    finally:
        nn.Module.cuda = saved


@contextlib.contextmanager
def injected_noise(noise_list):
    """Replace torch.rand_like by a FIFO of pre-drawn uniform tensors.

    LLongRangAttention.gumbel_softmax draws fresh uniform noise on every call
    (arch/SIDECVSR_our.py:2168-2171); parity needs the same draws on both
    sides.  Call order inside CVSR_V8.forward is neighbour i = 0,1,2,4,5,6
    (arch/SIDECVSR_our.py:4443-4452).
    """
    saved = torch.rand_like
    queue = list(noise_list)

    def fake(x, *a, **k):
        n = queue.pop(0)
        assert n.shape == x.shape, (n.shape, x.shape)
        return n.to(dtype=x.dtype, device=x.device)

    torch.rand_like = fake
    try:
        yield
    finally:
        torch.rand_like = saved


def build_reference_model(variant="O1"):
    """CVSR_V8 as the scripts build it (train_LD_37.py:20,165); variant 'O2'
    swaps in the DCN alignment that CVSR_V8 carries commented out
    (arch/SIDECVSR_our.py:4396)."""
    A = import_reference_arch()
    with cpu_construction():
        m = A.CVSR_V8()
        if variant == "O2":
            m.MV_deform_align = A.MVDualAttAlignment(
                64, 64, 3, padding=1, deformable_groups=16, max_residue_magnitude=10)
    return m.eval()
