"""cdfo_b200 -- B200 (sm_100a) hot path of CDFO's coding-prior-guided alignment, attention and upsampling tail.

The compute lives in libcdfo_b200.so (hand-written CUDA behind the C ABI of
include/cdfo_b200.h); this package is the thin host side that mirrors the
reference's operator / module interfaces.  There is no CPU fallback.
"""
from . import _lib, config, conv  # noqa: F401
from . import deform_conv_cuda  # noqa: F401
from .dcn import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack,  # noqa: F401
                  deform_conv, deform_conv2d, modulated_deform_conv)
from .attentionlayer import DSTA  # noqa: F401
from .priors import flow_warp, modify_mv_for_end_frames, mv2mvs, mv2mvs_ra  # noqa: F401
from .metrics import planes_to_unit, psnr_ssim, sr_to_u8  # noqa: F401

__version__ = "0.1.0"
