"""TEST INFRASTRUCTURE ONLY -- golden output of the REAL reference's DSTA block (ops/attentionlayer.py:86-156), produced in the build
container by running that file's own forward.  Run:  python -m oracle.make_golden_dsta

The reference's DSTA cannot run as shipped without a GPU: its ModulatedDeformConv calls the compiled extension `deform_conv_cuda`
(ops/dcn/deform_conv.py:11,144-148), which has no CPU path and refuses CPU tensors (:136-137).  Here
  * `deform_conv_cuda` is a stand-in module whose modulated_deform_conv_cuda_forward fills `output` with oracle/dcn_ref.c -- the plain C
    restatement of deform_conv_cuda_kernel.cu:467-496,570-632, pinned by the reference's own KAT (ops/dcn/simple_check.py) and by
    torchvision's CPU deform_conv2d (tests/test_oracle_dcn.py);
  * the input is a CPU tensor subclass that answers is_cuda = True, so the reference's Function.forward takes its CUDA branch.
Everything else -- the eleven convolutions, max-pool, two bilinear resizes, sigmoid gates, the wiring -- is the reference's code, unmodified.
Stored: tests/golden/dsta_golden.npz {"out": [2, 64, 40, 56] fp32}; inputs / weights are regenerated from seeds (dsta_inputs)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cdfo_b200 import synthetic  # noqa: E402
from oracle import c_oracle, ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def dsta_inputs(template):
    """Seeded weights (mask / mask2 scaled up: offsets of a few pixels) and input, shared with tests/."""
    sd = synthetic.seeded_state_dict(template, seed=9)
    for k in ("mask.weight", "mask2.weight"):
        sd[k] = sd[k] * 4.0
    sd["dcn.bias"] = torch.linspace(-0.2, 0.2, sd["dcn.bias"].numel())
    g = torch.Generator().manual_seed(77)
    return sd, torch.randn(2, 64, 40, 56, generator=g)


class _FakeCuda(torch.Tensor):
    is_cuda = property(lambda self: True)


def _standin_extension():
    m = types.ModuleType("deform_conv_cuda")

    def modulated_deform_conv_cuda_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w, stride_h, stride_w,
                                           pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
        n = lambda t: t.detach().as_subclass(torch.Tensor).numpy()  # noqa: E731
        y = c_oracle.dcn_forward(n(input), n(offset), n(mask), n(weight), n(bias) if with_bias else None, (stride_h, stride_w),
                                 (pad_h, pad_w), (dilation_h, dilation_w), group, deformable_group)
        output.as_subclass(torch.Tensor).copy_(torch.from_numpy(y))

    m.modulated_deform_conv_cuda_forward = modulated_deform_conv_cuda_forward
    return m


def main():
    if ref_import.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_import.REF_ROOT)
    for k in ("ops.dcn.deform_conv", "ops.dcn", "ops", "ops.attentionlayer"):
        sys.modules.pop(k, None)
    sys.modules["deform_conv_cuda"] = _standin_extension()
    import ops.attentionlayer as AL          # the reference file, unmodified
    m = AL.DSTA(64).eval()
    sd, x = dsta_inputs(m.state_dict())
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = m(x.as_subclass(_FakeCuda)).as_subclass(torch.Tensor)
    from oracle import torch_ref
    with torch.no_grad():
        port = torch_ref.dsta(sd, x)
    err = float((port - out).abs().max())
    print("oracle port vs reference DSTA: max|diff| = %.3g (max|out| %.3g)" % (err, float(out.abs().max())))
    assert err < 1e-5
    json_keys = {k: list(v.shape) for k, v in sd.items()}
    np.savez_compressed(os.path.join(GOLD, "dsta_golden.npz"), out=out.numpy(), keys=np.array(sorted(json_keys)),
                        shapes=np.array([str(json_keys[k]) for k in sorted(json_keys)]))
    print(os.path.getsize(os.path.join(GOLD, "dsta_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
