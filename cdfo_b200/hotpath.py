"""Execution of the hot-path modules of CVSR_V8 on the device.

Each function takes the parameter-holder module (cdfo_b200/model.py) plus activations and runs the path the
reference runs at the cited lines.  `backend` switches per stage between the hand-written kernels of
libcdfo_b200 ("cuda") and an interim cuDNN/ATen composition ("aten") that exists only for stages whose
fused kernel has not landed yet; DESIGN.md lists which stage is where.  Everything is CUDA-only.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib, config, conv, dcn_sm100
from .priors import flow_warp_chw

# stage -> "cuda" | "aten"
backend = {
    "flow_warp": "cuda",
    "dcn": "cuda",
}


def _lrelu(x):
    return F.leaky_relu(x, 0.1)


def _c(mod, x, **kw):
    return F.conv2d(x, mod.weight, mod.bias, **kw)


# ------------------------------------------------------------------------------------------ MDTA (A4 / A5 shared)
def _mdta(q, k, v, temperature, heads):
    b, c, h, w = q.shape
    sh = (b, heads, c // heads, h * w)
    q = F.normalize(q.reshape(sh), dim=-1)
    k = F.normalize(k.reshape(sh), dim=-1)
    attn = ((q @ k.transpose(-2, -1)) * temperature).softmax(dim=-1)
    return (attn @ v.reshape(sh)).reshape(b, c, h, w)


def _gate(seq, x):
    y = x.mean(dim=(2, 3), keepdim=True)
    y = F.relu(_c(seq._modules["0"], y))
    return torch.sigmoid(_c(seq._modules["2"], y))


def _warp(extra_feat, flow):
    if backend["flow_warp"] == "cuda":
        return flow_warp_chw(extra_feat.contiguous().float(), flow.contiguous().float())
    raise _lib.CdfoError("flow_warp has no other backend")


def _dual_mdta(mod, x, extra_feat, pred_feat, flow, relu_fused):
    """warp + fusion_out + the two MDTA passes + project_out (arch:3304-3337 / :3456-3490). Returns (o1, o2)."""
    warped = _warp(extra_feat, flow)
    fo = mod.fusion_out._modules["0"] if relu_fused else mod.fusion_out
    fused = F.conv2d(torch.cat([warped, pred_feat], 1), fo.weight)
    if relu_fused:
        fused = F.relu(fused)
    t, hd = mod.temperature, mod.num_heads
    o1 = F.conv2d(_mdta(x, fused, warped * _gate(mod.conv_du, warped), t, hd), mod.project_out.weight)
    o2 = F.conv2d(_mdta(x, fused, pred_feat * _gate(mod.conv_du, pred_feat), t, hd), mod.project_out.weight)
    return o1, o2


@torch.no_grad()
def dual_att_alignment(mod, x, extra_feat, pred_feat, flow):
    """DualAttAlignment.forward, arch:3455-3496 (flow [B,2,H,W])."""
    x = _expand_batch(x, extra_feat.size(0))
    o1, o2 = _dual_mdta(mod, x, extra_feat, pred_feat, flow, relu_fused=True)
    out = F.relu(F.conv2d(torch.cat([o1 + o2, x], 1), mod.fusion_out._modules["0"].weight))
    out = out * _gate(mod.CALayer.conv_du, out)
    for rb in (mod.ResidualBlock, mod.ResidualBlock1):
        out = out + _c(rb.conv2, F.relu(_c(rb.conv1, out, padding=1)), padding=1)
    return out + x


def _expand_batch(x, B):
    """x [xB, ...] shared by B = k * xB samples (sample b uses x[b % xB])."""
    return x if x.size(0) == B else x.repeat(B // x.size(0), 1, 1, 1)


_head_cache = {}


def _head_weights(c2, dg):
    """conv_offset[-1] with its output channels permuted into (dy_k, dx_k, m_k) triples, k = g*9 + tap
    (reference channels 2k, 2k+1, dg*18 + k: the chunk/cat of arch:3341-3345), packed for the tcgen05 conv."""
    key = id(c2.weight)
    hit = _head_cache.get(key)
    if hit is not None and hit[0] == (c2.weight._version, c2.bias._version):
        return hit[1], hit[2]
    k = torch.arange(dg * 9, device=c2.weight.device)
    perm = torch.stack([2 * k, 2 * k + 1, dg * 18 + k], dim=1).reshape(-1)
    w = c2.weight.detach().index_select(0, perm).contiguous()
    wpk = conv.pack_weight(w)
    bias = c2.bias.detach().index_select(0, perm).contiguous().float()
    _head_cache[key] = ((c2.weight._version, c2.bias._version), wpk, bias, w)
    return wpk, bias


@torch.no_grad()
def mv_offset_fields(mod, x, extra_feat, pred_feat, flow):
    """Learned offset residual and mask of MVDualAttAlignment WITHOUT the MV prior (arch:3339-3350 minus the
    `+ flow.flip(1).repeat(...)` term, which the DCN kernel adds itself), as packed fields
    [B, dg*9, H, W, 4] fp16 = (10*tanh(dy1) + 10*tanh(dy2), same for dx, sigmoid(m1 + m2), 0) per k = g*9 + tap.
    Both conv_offset layers run in the tcgen05 convolution kernel; tanh / sum / sigmoid are its epilogue."""
    B = extra_feat.size(0)
    o1, o2 = _dual_mdta(mod, _expand_batch(x, B), extra_feat, pred_feat, flow, relu_fused=False)
    c0, c2 = mod.conv_offset._modules["0"], mod.conv_offset._modules["2"]
    z = conv.conv3x3(conv.to_c8(torch.cat([o1, o2], 0)), c0.weight, c0.bias, conv.ACT_LRELU)   # [2B, 8, H, W, 8]
    H, W = z.shape[2:4]
    dg = mod.deformable_groups
    wpk, bias = _head_weights(c2, dg)
    first = torch.empty((B, dg * 9, H, W, 4), dtype=torch.float16, device=z.device)
    fields = torch.empty_like(first)
    args = (B, 64, dg, H, W, ctypes.c_float(float(mod.max_residue_magnitude)), _lib.stream_ptr(z.device))
    _lib.call("cdfo_mv_offset_head_sm100_fwd", _lib.ptr(z[:B]), _lib.ptr(wpk), _lib.ptr(bias), _lib.ptr(None),
              _lib.ptr(first), *args)
    _lib.call("cdfo_mv_offset_head_sm100_fwd", _lib.ptr(z[B:]), _lib.ptr(wpk), _lib.ptr(bias), _lib.ptr(first),
              _lib.ptr(fields), *args)
    return fields


def unpack_fields(fields):
    """Packed fields -> (residual [B, dg*18, H, W], mask [B, dg*9, H, W]) fp32 in the reference's channel order."""
    B, K, H, W, _ = fields.shape
    f = fields.float()
    residual = torch.stack([f[..., 0], f[..., 1]], dim=2).reshape(B, 2 * K, H, W)
    return residual, f[..., 2].contiguous()


@torch.no_grad()
def mv_dual_att_alignment(mod, x, extra_feat, pred_feat, flow):
    """MVDualAttAlignment.forward, arch:3303-3352: DCN with offset = residual + decoded MV prior.
    `x` may hold fewer samples than the other arguments (sample b uses x[b % x.size(0)]): the model's six
    neighbour calls share the centre-frame feature (arch:4456)."""
    fields = mv_offset_fields(mod, x, extra_feat, pred_feat, flow)
    if config.dcn_gather == "tex":
        return dcn_sm100.dcn_tex(dcn_sm100.pack_q4t(x), fields, dcn_sm100.pack_weight_f16(mod.weight), mod.bias, mv=flow)
    xq = dcn_sm100.pack_q4p(x)
    return dcn_sm100.dcn_sm100(xq, None, None, dcn_sm100.pack_weight(mod.weight), mod.bias, mv=flow, fused_fields=fields)


# ------------------------------------------------------------------------------------------ A8
_lra_tables = {}


def _lra_tap_tables(mod):
    """kw[9], kh[9], K1[64], R[64,64] of csrc/lra.cu from directW1_conv / directH1_conv (cached per weight version)."""
    key = (id(mod.directW1_conv.weight), mod.directW1_conv.weight._version, mod.directH1_conv.weight._version)
    hit = _lra_tables.get(id(mod))
    if hit is not None and hit[0] == key:
        return hit[1]
    kw = mod.directW1_conv.weight.detach().reshape(9).double()
    kh = mod.directH1_conv.weight.detach().reshape(9).double()
    c = torch.arange(64, device=kw.device)
    tix = c.view(64, 1) - c.view(1, 64) + 4                 # tix[c_star][c] = tap index of a bump centred at c_star
    tapm = torch.where((tix >= 0) & (tix <= 8), kw[tix.clamp(0, 8)], torch.zeros((), dtype=kw.dtype, device=kw.device))
    k1 = tapm.sum(dim=1)                                     # K1[c_star]
    r = tapm @ tapm.t()                                      # R[c1][c2]
    tab = torch.cat([kw, kh, k1, r.reshape(-1)]).float().contiguous()
    _lra_tables[id(mod)] = (key, tab)
    return tab


@torch.no_grad()
def long_range_attention(mod, res, x, u):
    """LLongRangAttention.forward, arch:2179-2249, u = uniform noise of gumbel_softmax (arch:2169).
    Mask logits (a global pooling of two small convolutions of the residual prior) and the 1x1 input_conv are
    ATen/cuDNN calls; the mask, the row / column / window attentions and the fuse convolution are csrc/lra.cu."""
    B, C, H, W = x.shape
    v = F.relu(_c(mod.conv_du_re._modules["0"], res))
    v = F.relu(_c(mod.conv_du_re._modules["2"], v, stride=2, padding=2))
    v = v.mean(dim=(2, 3), keepdim=True)
    vmax = F.relu(_c(mod.conv_du_re2._modules["0"], v)).reshape(B, C).contiguous()   # bilinear up of a 1x1 map = broadcast
    x = x.contiguous()
    qv = _c(mod.input_conv, x).contiguous()
    out = torch.empty_like(x)
    nbytes = _lib.lib().cdfo_lra_workspace_bytes(B, H, W)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _lib.call("cdfo_lra_fwd", _lib.ptr(qv), _lib.ptr(u.contiguous()), _lib.ptr(vmax), _lib.ptr(x), _lib.ptr(_lra_tap_tables(mod)),
              ctypes.c_float(float(mod.directW1_conv.bias)), ctypes.c_float(float(mod.directH1_conv.bias)),
              _lib.ptr(mod.fuse.weight.detach().reshape(64, 128).contiguous()), _lib.ptr(mod.fuse.bias.detach().contiguous()),
              _lib.ptr(out), _lib.ptr(ws), B, H, W, _lib.stream_ptr(x.device))
    return out


# ------------------------------------------------------------------------------------------ model-level stages
@torch.no_grad()
def align_neighbours(model, center, fea_nb, ufs_nb, rms_nb, mv_nb, u_nb):
    """The body of the reference's neighbour loop (arch:4445-4456) for all six neighbours at once.
    Batch index = n * B + b (neighbour-major); `center` [B,64,H,W] is shared by the six."""
    ufs_prior = _c(model.conv_expand_ufs, ufs_nb, padding=1)
    rms_prior = _c(model.conv_expand_rms, rms_nb, padding=1)
    x_n = long_range_attention(model.RDAB, rms_prior, fea_nb + rms_prior, u_nb)
    fr = model.conv_expand_fea_r
    fea_i = conv.conv3x3(conv.to_c8(torch.cat([fea_nb, x_n], 1)), fr.weight, fr.bias, conv.ACT_NONE, out_nchw=True)
    return model.MV_deform_align(center, fea_i, ufs_prior, mv_nb)


@torch.no_grad()
def temporal_fusion(model, aligned, center, B):
    """stack + tsa_fusion 1x1 + lrelu, arch:4463-4466. aligned [6B,64,H,W] neighbour-major."""
    _, c, h, w = aligned.shape
    a = aligned.view(6, B, c, h, w)
    stacked = torch.cat([a[0], a[1], a[2], center, a[3], a[4], a[5]], dim=1)   # [B, 448, H, W], frame order 0..6
    return _lrelu(_c(model.tsa_fusion, stacked))


@torch.no_grad()
def tail(model, t, x_center):
    """arch:4473-4480."""
    out = _lrelu(F.pixel_shuffle(_c(model.upconv1, t), 2))
    out = _lrelu(F.pixel_shuffle(_c(model.upconv2, out), 2))
    out = _c(model.conv_last, out, padding=1)
    return out + F.interpolate(x_center, scale_factor=4.0, mode="bilinear", align_corners=False)
