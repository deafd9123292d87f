"""Run-to-run determinism of the benchmarked step (272x480, B = 2, bf16, feature ring): the same two frames twice in one process, with the
caching allocator's free blocks poisoned with NaNs in between (a kernel that reads memory it did not write shows up as a difference or a
NaN), stage by stage.  Prints one line per stage; exit code 1 on any difference."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cdfo_b200  # noqa: E402
from cdfo_b200 import hotpath, synthetic  # noqa: E402
from cdfo_b200.model import CVSR_V8  # noqa: E402


def poison(dev):
    free, _ = torch.cuda.mem_get_info(dev)
    blocks = []
    try:
        for _ in range(6):
            blocks.append(torch.full((256 << 20,), float("nan"), dtype=torch.float32, device=dev))      # 1 GiB each
    except RuntimeError:
        pass
    del blocks


def main():
    dev = torch.device("cuda:0")
    H, W, B = 272, 480, 2
    m = CVSR_V8(alignment="mv_dcn")
    m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), seed=4), strict=True)
    m = m.to(dev).eval()
    m.lowp = torch.bfloat16
    m.feature_ring = True
    clip = {k: v.to(dev) for k, v in synthetic.make_clip(77, H, W, B).items()}
    mvs = torch.cat([cdfo_b200.mv2mvs(clip["mv_l0"][s]) for s in range(B)], 0)
    noise = torch.cat([u.to(dev) for u in synthetic.gumbel_uniforms(4, 9, 0, B, H, W)], 0)
    taps = {}
    orig = {}

    def wrap(name, fn):
        def f(*a, **k):
            out = fn(*a, **k)
            t = out[0] if isinstance(out, tuple) else out
            if torch.is_tensor(t):
                taps.setdefault(name, []).append(t.detach().float().clone())
            return out
        return f
    for name in ("long_range_attention", "dual_mdta", "mv_hidden_maps", "align_and_fuse", "recon_trunk", "tail", "prior_conv"):
        orig[name] = getattr(hotpath, name)
        setattr(hotpath, name, wrap(name, orig[name]))
    runs = []
    for r in range(2):
        taps.clear()
        sr0, l1 = m(clip["x"], None, mvs, clip["pms"], clip["rms"], clip["ufs"], None, noise=noise)
        sr1, l1 = m(clip["x"], None, mvs, clip["pms"], clip["rms"], clip["ufs"], l1, noise=noise)
        torch.cuda.synchronize()
        runs.append(({k: [t.cpu() for t in v] for k, v in taps.items()}, sr0.float().cpu(), sr1.float().cpu()))
        del sr0, sr1, l1
        poison(dev)
    bad = 0
    (ta, a0, a1), (tb, b0, b1) = runs
    for name in ta:
        for i, (x, y) in enumerate(zip(ta[name], tb[name])):
            same = torch.equal(x, y)
            nan = bool(torch.isnan(x).any() or torch.isnan(y).any())
            if not same or nan:
                bad += 1
            print("%-22s call %2d  identical=%s  nan=%s  max|diff|=%.3g" % (name, i, same, nan, float((x - y).abs().max()) if not nan else float("nan")))
    for name, x, y in (("SR first frame", a0, b0), ("SR cached frame", a1, b1)):
        same = torch.equal(x, y)
        bad += 0 if same else 1
        print("%-22s          identical=%s  max|diff|=%.3g" % (name, same, float((x - y).abs().max())))
    print("determinism probe:", "OK" if bad == 0 else "%d stage outputs differ" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
