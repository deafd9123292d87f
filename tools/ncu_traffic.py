"""Writes profiles/rNN_roofline_kernel_traffic.json -- the measured DRAM traffic of the roofline kernel that bench.py reports as
`roofline.traffic` -- from an `ncu --set full` report:
    python tools/ncu_traffic.py gpurun_out/fused.ncu-rep r02 mv_head_dcn_fused <pixels per launch>
(dram__bytes_read.sum + dram__bytes_write.sum of the launches whose kernel name contains the pattern, averaged)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(rep, rnd, pattern, px):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    rd, wr, dur, names = [], [], [], []
    for r in rows[2:]:
        if pattern not in r[col["Kernel Name"]]:
            continue
        names.append(r[col["Kernel Name"]])
        rd.append(float(r[col["dram__bytes_read.sum"]].replace(",", "")) * SCALE[units[col["dram__bytes_read.sum"]]])
        wr.append(float(r[col["dram__bytes_write.sum"]].replace(",", "")) * SCALE[units[col["dram__bytes_write.sum"]]])
        dur.append(r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]])
    if not rd:
        raise SystemExit("no launch matching %r in %s" % (pattern, rep))
    out = {"round": rnd, "source": os.path.relpath(rep, ROOT), "kernel": names[0], "launches": len(rd),
           "dram_bytes_read": sum(rd) / len(rd), "dram_bytes_write": sum(wr) / len(wr), "px_per_launch": int(px),
           "dram_bytes_per_px": (sum(rd) + sum(wr)) / len(rd) / float(px), "gpu_time_under_ncu": dur}
    path = os.path.join(ROOT, "profiles", "%s_roofline_kernel_traffic.json" % rnd)
    json.dump(out, open(path, "w"), indent=1)
    print(path, out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4])
