"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz|json by importing and running the REAL reference
(/root/reference, read-only) in the build container.  Run:  python -m oracle.make_golden

What is recorded (inputs are NOT stored: every test regenerates them from the same seeds through
cdfo_b200/synthetic.py; weights likewise through synthetic.seeded_state_dict):

  state_dict_{O1,O2}.json   parameter names + shapes of the reference models (drop-in contract)
  priors_golden.npz         mv2mvs / modify_mv_for_end_frames / generate_input_index outputs of the reference's
                            own functions (extracted from test_LD_37.py with ast; the script cannot be imported)
  modules_golden.npz        outputs of the reference's hot-path modules on seeded feature-level inputs
                            (flow_warp, LLongRangAttention, DualAttAlignment, MVDualAttAlignment incl. offset /
                            mask fields, tail), B=2, 24x40
  model_golden.npz          CVSR_V8 (as shipped, "O1") and CVSR_V8 with MVDualAttAlignment ("O2"):
                            SR output of a first frame and of a second frame through the L1_fea cache, 64x64 LR
"""
import ast
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cdfo_b200 import synthetic  # noqa: E402  (pure torch/numpy helpers: seeds -> inputs / weights)
from oracle import priors_ref, ref_import, torch_ref  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MOD_B, MOD_H, MOD_W = 2, 24, 40


def module_inputs(seed=11, B=MOD_B, H=MOD_H, W=MOD_W):
    """Feature-level inputs of the hot-path modules (shared with tests/)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    d = {
        "x": r(B, 64, H, W) * 0.5, "extra": r(B, 64, H, W) * 0.5, "pred": r(B, 64, H, W) * 0.5,
        "res": r(B, 64, H, W) * 0.3, "u": torch.rand(B, 64, H, W, generator=g).clamp_min(1e-12),
        "trunk_out": r(B, 64, H, W) * 0.3, "x_center": torch.rand(B, 1, H, W, generator=g),
    }
    mvq = torch.randint(-192, 192, (B, 2, H // 8, W // 8), generator=g).float() / 128.0
    d["flow"] = mvq.repeat_interleave(8, 2).repeat_interleave(8, 3).contiguous()
    d["flow"][:, :, 0, 0] = 50.0  # one far-out-of-frame vector
    return d


def frame_inputs(seed, H=64, W=64, B=1):
    clip = synthetic.make_clip(seed, H, W, B)
    mvs = torch.stack([torch.from_numpy(priors_ref.mv2mvs_model_layout(clip["mv_l0"][b].numpy())[0]) for b in range(B)])
    return clip, mvs


def extract_reference_functions():
    src = open(os.path.join(ref_import.REF_ROOT, "test_LD_37.py")).read()
    ns = {"np": np, "torch": torch}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("mv2mvs", "modify_mv_for_end_frames", "generate_input_index"):
            exec(compile(ast.Module([node], []), "test_LD_37.py", "exec"), ns)
    return ns


def priors_cases():
    rng = np.random.default_rng(5)
    mv = rng.integers(-128, 128, (16, 24, 3)).astype(np.int8)
    mv[..., 2] = rng.choice([-1, -2, -4, 1, 3], (16, 24))
    mv[:2, :, 2] = 0
    mv[:1, :12, :2] = 0
    return mv, [(0, 10), (1, 10), (2, 10), (9, 10), (8, 10), (7, 10), (5, 10), (0, 3), (1, 3), (2, 3), (0, 1)]


def main():
    os.makedirs(GOLD, exist_ok=True)
    warnings.simplefilter("ignore")
    torch.set_num_threads(os.cpu_count())
    A = ref_import.import_reference_arch()

    # ---------------- priors
    ns = extract_reference_functions()
    mv, fix_cases = priors_cases()
    out = {"mv2mvs": ns["mv2mvs"](mv.copy()).numpy()}
    base = torch.from_numpy(priors_ref.mv2mvs_model_layout(mv))
    for i, mx in fix_cases:
        t = base.clone()
        ns["modify_mv_for_end_frames"](i, t, mx)
        out["fix_%d_%d" % (i, mx)] = t.numpy()
    out["index_5_7_31"] = ns["generate_input_index"](5, 7, 31)
    out["index_0_7_31"] = ns["generate_input_index"](0, 7, 31)
    out["index_30_7_31"] = ns["generate_input_index"](30, 7, 31)
    np.savez_compressed(os.path.join(GOLD, "priors_golden.npz"), **out)

    # ---------------- models + state-dict contract
    models = {}
    for variant in ("O1", "O2"):
        m = ref_import.build_reference_model(variant)
        tmpl = m.state_dict()
        json.dump({k: list(v.shape) for k, v in tmpl.items()},
                  open(os.path.join(GOLD, "state_dict_%s.json" % variant), "w"), indent=0, sort_keys=True)
        m.load_state_dict(synthetic.seeded_state_dict(tmpl, seed=4), strict=True)
        models[variant] = m

    # ---------------- module-level goldens (B=2, 24x40)
    d = module_inputs()
    mod = {}
    with torch.no_grad():
        mod["flow_warp"] = A.flow_warp(d["extra"], d["flow"].permute(0, 2, 3, 1)).numpy()
        m1, m2 = models["O1"], models["O2"]
        with ref_import.injected_noise([d["u"]]):
            mod["lra_out"] = m1.RDAB(d["res"], d["x"]).numpy()
        sd1 = {k: v for k, v in m1.state_dict().items()}
        mod["lra_mask_bits"] = np.packbits(torch_ref.lra_mask(sd1, "RDAB.", d["res"], d["u"]).numpy().astype(np.uint8))
        mod["dual_att_out"] = m1.MV_deform_align(d["x"], d["extra"], d["pred"], d["flow"]).numpy()
        mod["mv_dcn_out"] = m2.MV_deform_align(d["x"], d["extra"], d["pred"], d["flow"]).numpy()
        sd2 = {k: v for k, v in m2.state_dict().items()}
        off, msk = torch_ref.mv_offsets(sd2, "MV_deform_align.", d["x"], d["extra"], d["pred"], d["flow"])
        mod["mv_offset_g0g15"] = torch.cat([off[:, :18], off[:, -18:]], 1).numpy()   # deformable groups 0 and 15
        mod["mv_mask_g0g15"] = torch.cat([msk[:, :9], msk[:, -9:]], 1).numpy()
        t = m1.lrelu(m1.pixel_shuffle(m1.upconv1(d["trunk_out"].clone())))
        t = m1.lrelu(m1.pixel_shuffle(m1.upconv2(t)))
        t = m1.conv_last(t)
        mod["tail_out"] = (t + torch.nn.functional.interpolate(d["x_center"], scale_factor=4.0, mode="bilinear",
                                                               align_corners=False)).numpy()
        # the oracle restatement must agree with the reference before it is trusted anywhere else
        chk = {
            "lra_out": torch_ref.long_range_attention(sd1, "RDAB.", d["res"], d["x"], d["u"]),
            "dual_att_out": torch_ref.dual_att_alignment(sd1, "MV_deform_align.", d["x"], d["extra"], d["pred"], d["flow"]),
            "mv_dcn_out": torch_ref.mv_dual_att_alignment(sd2, "MV_deform_align.", d["x"], d["extra"], d["pred"], d["flow"]),
            "tail_out": torch_ref.tail(sd1, d["trunk_out"], d["x_center"]),
            "flow_warp": torch_ref.flow_warp(d["extra"], d["flow"].permute(0, 2, 3, 1)),
        }
        for k, v in chk.items():
            err = float(np.abs(v.numpy() - mod[k]).max())
            print("oracle vs reference  %-14s max|diff| = %.3g" % (k, err))
            assert err < 2e-5, k
    np.savez_compressed(os.path.join(GOLD, "modules_golden.npz"), **mod)

    # ---------------- full-model goldens (64x64 LR, first frame + cached second frame)
    full = {}
    clip0, mvs0 = frame_inputs(1)
    clip1, mvs1 = frame_inputs(2)
    # second window = first window shifted by one frame + one new frame (what the L1_fea cache assumes)
    for k in ("x", "pms", "rms", "ufs"):
        clip1[k] = torch.cat([clip0[k][:, 1:], clip1[k][:, -1:]], 1)
    n0 = synthetic.gumbel_uniforms(4, 0, 0, 1, 64, 64)
    n1 = synthetic.gumbel_uniforms(4, 0, 1, 1, 64, 64)
    with torch.no_grad():
        for variant, m in models.items():
            sd = dict(m.state_dict())
            with ref_import.injected_noise(n0):
                sr0, l1 = m(clip0["x"], mvs0, mvs0, clip0["pms"], clip0["rms"], clip0["ufs"])
            with ref_import.injected_noise(n1):
                sr1, l1b = m(clip1["x"], mvs1, mvs1, clip1["pms"], clip1["rms"], clip1["ufs"], l1)
            full["%s_sr0" % variant] = sr0.numpy()
            full["%s_sr1" % variant] = sr1.numpy()
            full["%s_l1_mean" % variant] = l1.mean(dim=(2, 3)).numpy()
            full["%s_l1_std" % variant] = l1.std(dim=(2, 3)).numpy()
            o0, ol1 = torch_ref.cvsr_v8_forward(sd, clip0["x"], mvs0, clip0["pms"], clip0["rms"], clip0["ufs"], None, n0, variant)
            o1, _ = torch_ref.cvsr_v8_forward(sd, clip1["x"], mvs1, clip1["pms"], clip1["rms"], clip1["ufs"], ol1, n1, variant)
            e0, e1 = float((o0 - sr0).abs().max()), float((o1 - sr1).abs().max())
            print("oracle vs reference  CVSR_V8 %s  frame0 %.3g  frame1(cached) %.3g   |sr| range [%.3f, %.3f]"
                  % (variant, e0, e1, float(sr0.min()), float(sr0.max())))
            assert e0 < 1e-4 and e1 < 1e-4
    np.savez_compressed(os.path.join(GOLD, "model_golden.npz"), **full)
    for f in sorted(os.listdir(GOLD)):
        print("%8d  %s" % (os.path.getsize(os.path.join(GOLD, f)), f))


if __name__ == "__main__":
    main()
