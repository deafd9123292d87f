"""Which ATen / cuDNN calls are left in one steady-state step, and where they come from (torch.profiler with Python stacks).

Prints device time per (op, innermost cdfo_b200 source line); kernels launched through the C ABI have no ATen op and
show up only in the total.  Run on the GPU box: python tools/prof_glue.py [--seqs 2]
"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdfo_b200 import synthetic  # noqa: E402
from cdfo_b200.model import CVSR_V8  # noqa: E402
import cdfo_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=2)
    ap.add_argument("--H", type=int, default=272)
    ap.add_argument("--W", type=int, default=480)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    S, H, W = a.seqs, a.H, a.W
    m = CVSR_V8(alignment="mv_dcn")
    m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), 4))
    m = m.to(dev).eval()
    m.lowp = torch.bfloat16
    clip = synthetic.make_clip(0, H, W, S)
    d = {k: clip[k].to(dev) for k in ("x", "pms", "rms", "ufs")}
    mvs = torch.cat([cdfo_b200.mv2mvs(clip["mv_l0"][s].to(dev)) for s in range(S)], 0)
    g = torch.Generator(device=dev).manual_seed(1)
    noise = [torch.rand((S, 64, H, W), device=dev, generator=g).clamp_min_(1e-12) for _ in range(6)]
    _, l1 = m(d["x"], None, mvs, d["pms"], d["rms"], d["ufs"], None, noise=noise)
    for _ in range(3):
        _, l1 = m(d["x"], None, mvs, d["pms"], d["rms"], d["ufs"], l1, noise=noise)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        _, l1 = m(d["x"], None, mvs, d["pms"], d["rms"], d["ufs"], l1, noise=noise)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    total = 0.0
    for ev in prof.events():
        dt = getattr(ev, "self_device_time_total", 0.0)
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            total += ev.device_time_total if hasattr(ev, "device_time_total") else 0.0
            continue
        if dt <= 0:
            continue
        where = next((s for s in (ev.stack or []) if "cdfo_b200/" in s), "?")
        where = where.split("cdfo_b200/")[-1]
        e = agg[(ev.name, where)]
        e[0] += dt
        e[1] += 1
    aten = sum(v[0] for v in agg.values())
    print("device time in kernels: %.2f ms; in ATen / cuDNN ops: %.2f ms" % (total / 1e3, aten / 1e3))
    for (name, where), (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:70]:
        print("%8.1f us %4d  %-34s %s" % (t, c, name[:34], where[:90]))


if __name__ == "__main__":
    main()
