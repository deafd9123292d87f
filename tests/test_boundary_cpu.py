"""Drop-in boundary, host side (no GPU): the torch.library override of torchvision::deform_conv2d registers / unregisters cleanly, and
cdfo_b200.deform_conv_cuda mounted under the reference's OWN ops/dcn/deform_conv.py receives exactly the positional arguments the
reference passes (ops/dcn/deform_conv.py:52-57,144-148 -> pybind signatures deform_conv_cuda.cpp:151-156,486-492)."""
import sys

import pytest
import torch

from oracle import ref_import


def test_torchvision_override_registers_cuda_key_only():
    import torchvision
    from cdfo_b200 import torchvision_override as tvo
    assert tvo.install() and tvo.installed()
    try:
        dump = torch._C._dispatch_dump("torchvision::deform_conv2d")
        cuda = [ln for ln in dump.splitlines() if ln.startswith("CUDA:")]
        assert len(cuda) == 1 and "torchvision_override.py" in cuda[0]
        assert any(ln.startswith("CUDA (inactive):") and "deform_conv2d_kernel.cu" in ln for ln in dump.splitlines())
        # the CPU kernel is untouched: zero offsets -> plain convolution
        x, off, w = torch.randn(1, 4, 6, 6), torch.zeros(1, 18, 6, 6), torch.randn(4, 4, 3, 3)
        y = torchvision.ops.deform_conv2d(x, off, w, None, 1, 1, 1)
        assert (y - torch.nn.functional.conv2d(x, w, padding=1)).abs().max().item() < 1e-5
    finally:
        tvo.uninstall()
    dump = torch._C._dispatch_dump("torchvision::deform_conv2d")
    assert not any("torchvision_override.py" in ln for ln in dump.splitlines()) and not tvo.installed()


class _FakeCuda(torch.Tensor):
    """A CPU tensor that answers is_cuda = True, so that the reference's Function.forward takes its CUDA branch here."""
    is_cuda = property(lambda self: True)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_module_mounts_under_the_reference_deform_conv_py(monkeypatch):
    import cdfo_b200.deform_conv_cuda as ours
    saved = {k: sys.modules.get(k) for k in ("deform_conv_cuda", "ops.dcn.deform_conv", "ops.dcn", "ops")}
    try:
        for k in ("ops.dcn.deform_conv", "ops.dcn", "ops"):
            sys.modules.pop(k, None)
        sys.modules["deform_conv_cuda"] = ours
        if ref_import.REF_ROOT not in sys.path:
            sys.path.insert(0, ref_import.REF_ROOT)
        import ops.dcn.deform_conv as dc                 # the reference file, unmodified
        assert dc.deform_conv_cuda is ours
        seen = {}

        def rec_v2(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                   dilation_h, dilation_w, group, deformable_group, with_bias):
            seen["v2"] = dict(x=tuple(input.shape), w=tuple(weight.shape), b=tuple(bias.shape), off=tuple(offset.shape), m=tuple(mask.shape),
                              out=tuple(output.shape), scratch=(ones.numel(), columns.numel()), k=(kernel_h, kernel_w), s=(stride_h, stride_w),
                              p=(pad_h, pad_w), d=(dilation_h, dilation_w), g=group, dg=deformable_group, with_bias=with_bias)
            output.zero_()

        def rec_v1(input, weight, offset, output, columns, ones, kW, kH, dW, dH, padW, padH, dilationW, dilationH, group, deformable_group,
                   im2col_step):
            seen["v1"] = dict(x=tuple(input.shape), w=tuple(weight.shape), off=tuple(offset.shape), out=tuple(output.shape), k=(kH, kW),
                              s=(dH, dW), p=(padH, padW), d=(dilationH, dilationW), g=group, dg=deformable_group, step=im2col_step)
            output.zero_()
            return 1

        monkeypatch.setattr(ours, "modulated_deform_conv_cuda_forward", rec_v2)
        monkeypatch.setattr(ours, "deform_conv_forward_cuda", rec_v1)
        fc = lambda *s: torch.zeros(*s).as_subclass(_FakeCuda)  # noqa: E731
        m = dc.ModulatedDeformConv(8, 6, 3, stride=1, padding=1, dilation=1, groups=1, deformable_groups=2, bias=True)
        with torch.no_grad():
            y = m(fc(2, 8, 10, 12), fc(2, 2 * 2 * 9, 10, 12), fc(2, 2 * 9, 10, 12))
        assert tuple(y.shape) == (2, 6, 10, 12)
        assert seen["v2"] == dict(x=(2, 8, 10, 12), w=(6, 8, 3, 3), b=(6,), off=(2, 36, 10, 12), m=(2, 18, 10, 12), out=(2, 6, 10, 12),
                                  scratch=(0, 0), k=(3, 3), s=(1, 1), p=(1, 1), d=(1, 1), g=1, dg=2, with_bias=True)
        m1 = dc.DeformConv(8, 6, 3, stride=1, padding=1, deformable_groups=2)
        with torch.no_grad():
            y1 = m1(fc(2, 8, 10, 12), fc(2, 36, 10, 12))
        assert tuple(y1.shape) == (2, 6, 10, 12)
        assert seen["v1"] == dict(x=(2, 8, 10, 12), w=(6, 8, 3, 3), off=(2, 36, 10, 12), out=(2, 6, 10, 12), k=(3, 3), s=(1, 1), p=(1, 1),
                                  d=(1, 1), g=1, dg=2, step=2)
        # un-patched: the real adapter refuses CPU tensors the way the reference's own op would not even be reached
        monkeypatch.undo()
        with pytest.raises(NotImplementedError):
            dc.modulated_deform_conv(torch.zeros(1, 8, 4, 4), torch.zeros(1, 36, 4, 4), torch.zeros(1, 18, 4, 4), torch.zeros(6, 8, 3, 3),
                                     None, 1, 1, 1, 1, 2)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
