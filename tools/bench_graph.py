"""Eager vs CUDA-graph replay of the steady-state step at c3 (2 sequences): same outputs, time per step."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cdfo_b200  # noqa: E402
from cdfo_b200 import synthetic  # noqa: E402
from cdfo_b200.graph import GraphedStep  # noqa: E402
from cdfo_b200.model import CVSR_V8  # noqa: E402

dev = torch.device("cuda:0")
S, H, W = 2, 272, 480
m = CVSR_V8(alignment="mv_dcn")
m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), 4))
m = m.to(dev).eval(); m.lowp = torch.bfloat16
clip = synthetic.make_clip(0, H, W, S)
d = {k: clip[k].to(dev) for k in ("x", "pms", "rms", "ufs")}
mvs = torch.cat([cdfo_b200.mv2mvs(clip["mv_l0"][s].to(dev)) for s in range(S)], 0)
g = torch.Generator(device=dev).manual_seed(1)
noise = [torch.rand((S, 64, H, W), device=dev, generator=g).clamp_min_(1e-12) for _ in range(6)]
_, l1 = m(d["x"], None, mvs, d["pms"], d["rms"], d["ufs"], None, noise=noise)


def timeit(fn, iters=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


sr_e, l1_e = m(d["x"], None, mvs, d["pms"], d["rms"], d["ufs"], l1, noise=noise)
t_eager = timeit(lambda: m(d["x"], None, mvs, d["pms"], d["rms"], d["ufs"], l1, noise=noise))
gs = GraphedStep(m, d["x"], mvs, d["pms"], d["rms"], d["ufs"], l1, noise)
sr_g, l1_g = gs(d["x"], mvs, d["pms"], d["rms"], d["ufs"], l1, noise)
torch.cuda.synchronize()
same = bool(torch.equal(sr_g, sr_e)) and bool(torch.equal(l1_g, l1_e))
t_graph = timeit(lambda: gs(d["x"], mvs, d["pms"], d["rms"], d["ufs"], l1, noise))
print(json.dumps({"eager_ms": round(t_eager, 3), "graph_ms": round(t_graph, 3), "identical_outputs": same,
                  "max_abs_diff": float((sr_g - sr_e).abs().max())}))
