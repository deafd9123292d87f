"""Kernel-level time split of the feature extraction for one new frame of S sequences at 272x480."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdfo_b200 import synthetic  # noqa: E402
from cdfo_b200.model import CVSR_V8  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    S, H, W = 2, 272, 480
    m = CVSR_V8(alignment="mv_dcn")
    m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), seed=4), strict=True)
    m = m.to(dev).eval()
    m.lowp = torch.bfloat16
    x = torch.rand(S, 1, H, W, device=dev)
    pm = torch.rand(S, 1, H, W, device=dev)
    with torch.no_grad():
        for _ in range(3):
            m._features(x, pm)
        torch.cuda.synchronize()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                m._features(x, pm)
            torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print("total device time per call: %.3f ms" % (tot / 5 / 1e3))
    for e in rows[:28]:
        print("%8.1f us  x%-4d %s" % (e.device_time_total / 5, e.count // 5, e.key[:110]))


if __name__ == "__main__":
    main()
