"""Interface mirror of the reference's spatial-attention block `DSTA` (ops/attentionlayer.py:86-156).

No model or script of the reference instantiates DSTA, so this is the boundary only: same constructor,
same parameter names / shapes (reference state_dicts load with strict=True), same forward(x) -> x * m * y.  Its internal
deformable convolution (16 -> 16 channels, 3x3, one deformable group per channel, on the max-pooled map) goes through
`cdfo_b200.ModulatedDeformConv`, i.e. the C-ABI kernel `cdfo_dcn_fwd`; the small dense convolutions around it are
cuDNN calls (the pooled map is 1/36 of the frame).  CUDA only, like the reference's op (ops/dcn/deform_conv.py:136-137).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .dcn import ModulatedDeformConv


class DSTA(nn.Module):
    def __init__(self, n_feats):
        super().__init__()
        f = n_feats // 4
        self.f = f
        self.conv1 = nn.Conv2d(n_feats, f, 1)                     # ops/attentionlayer.py:90-96
        self.conv_f = nn.Conv2d(f, f, 1)
        self.conv_max = nn.Conv2d(f, f, 3, padding=1)
        self.conv2 = nn.Conv2d(f, f, 3, stride=2, padding=0)
        self.conv3 = nn.Conv2d(f, f, 3, padding=1)
        self.conv3_ = nn.Conv2d(f, f, 3, padding=1)
        self.conv4 = nn.Conv2d(f, n_feats, 1)
        self.dcn = ModulatedDeformConv(f, f, 3, padding=1, deformable_groups=f)   # :100
        self.mask = nn.Conv2d(f, f * 27, 3, padding=1)            # :101
        self.down_conv2 = nn.Sequential(nn.Conv2d(f, f, 3, stride=2, padding=1), nn.ReLU(inplace=True))   # :104-106
        self.mask2 = nn.Conv2d(f, f * 27, 3, padding=1)           # :107
        self.conv_du = nn.Sequential(nn.Conv2d(f, 2 * f, 1), nn.ReLU(inplace=True), nn.Conv2d(2 * f, n_feats, 1),
                                     nn.Sigmoid())                # :110-115

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise NotImplementedError("cdfo_b200.DSTA is CUDA-only (its deformable convolution has no CPU path)")
        f = self.f
        reduced = self.conv1(x)                                                        # :119
        pooled = F.max_pool2d(self.conv2(reduced), kernel_size=7, stride=3)            # :120-121
        t = F.relu(self.conv3_(F.relu(self.conv3(F.relu(self.conv_max(pooled))))))     # :122-125
        coarse = self.mask2(self.down_conv2(t))                                        # :126-127
        fields = self.mask(t) + F.interpolate(coarse, t.shape[2:], mode="bilinear", align_corners=False)   # :128-130
        offset, msk = fields[:, :18 * f].contiguous(), torch.sigmoid(fields[:, 18 * f:]).contiguous()       # :131-134
        d = F.relu(self.dcn(pooled.contiguous(), offset, msk))                         # :135-136
        channel_weight = self.conv_du(F.adaptive_avg_pool2d(d, 1))                     # :137-138
        d = F.interpolate(d, x.shape[2:], mode="bilinear", align_corners=False)        # :139
        spatial = torch.sigmoid(self.conv4(d + self.conv_f(reduced)))                  # :140-142
        return x * spatial * channel_weight                                            # :156
