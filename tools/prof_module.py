"""torch.profiler kernel table of one hot-path stage at c3 (which kernels the stage's time is made of)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdfo_b200  # noqa: E402
from cdfo_b200 import hotpath, synthetic  # noqa: E402
from cdfo_b200.model import CVSR_V8  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--stage", default="rdab")
ap.add_argument("--seqs", type=int, default=2)
ap.add_argument("--plain", action="store_true", help="run the stage three times without torch.profiler (for ncu)")
a = ap.parse_args()
dev = torch.device("cuda:0")
S, H, W = a.seqs, 272, 480
m = CVSR_V8(alignment="mv_dcn")
m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), 4))
m = m.to(dev).eval(); m.lowp = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
fea_nb, center = r(6 * S, 64, H, W), r(S, 64, H, W)
rp, up = r(6 * S, 64, H, W) * 0.1, r(6 * S, 64, H, W) * 0.3
u = torch.rand(6 * S, 64, H, W, device=dev, generator=g).clamp_min(1e-12)
mv = (torch.randint(-192, 192, (6 * S, 2, H // 8, W // 8), device=dev, generator=g).float() / 128).repeat_interleave(8, 2).repeat_interleave(8, 3)
x1 = torch.rand(S, 1, H, W, device=dev, generator=g)
fns = {
    "rdab": lambda: hotpath.long_range_attention(m.RDAB, rp, fea_nb + rp, u),
    "align": lambda: m.MV_deform_align(center, fea_nb, up, mv),
    "trunk": lambda: m._trunk(cdfo_b200.conv.to_c8(center)),
    "tail": lambda: hotpath.tail(m, center, x1),
    "features": lambda: m._features(x1, x1),
}
if a.stage == "step":
    # one steady-state step of the whole model (cached L1 features, feature ring), as bench.py runs it
    m.feature_ring = True
    clip = synthetic.make_clip(1, H, W, S)
    d = {k: clip[k].to(dev) for k in ("x", "pms", "rms", "ufs")}
    mvs = torch.cat([cdfo_b200.mv2mvs(clip["mv_l0"].to(dev)[s]) for s in range(S)], 0)
    noise = torch.rand(6 * S, 64, H, W, device=dev, generator=g).clamp_min(1e-12)
    state = {}
    with torch.no_grad():
        _, state["l1"] = m(d["x"], None, mvs, d["pms"], d["rms"], d["ufs"], None, noise=noise)

    def step():
        _, state["l1"] = m(d["x"], None, mvs, d["pms"], d["rms"], d["ufs"], state["l1"], noise=noise)
    fns["step"] = step
fn = fns[a.stage]
with torch.no_grad():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if a.plain:
        fn()
        torch.cuda.synchronize()
        sys.exit(0)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
