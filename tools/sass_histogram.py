"""SASS evidence per kernel of libcdfo_b200.so: counts of the Blackwell-native opcodes (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UBLKCP = TMA, UTCBAR = tcgen05.commit, TEX) and of the legacy tensor path (HMMA = mma.sync) -> profiles/rNN_sass_opcodes.md.
    python tools/sass_histogram.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMASTG", "TEX", "HMMA", "MUFU", "SYNCS", "LDGSTS"]


def main(rnd):
    so = os.path.join(ROOT, "cdfo_b200", "libcdfo_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k + ("." + ".".join(op.split(".")[1:3]) if k in ("UTCHMMA", "UTMALDG", "HMMA") and "." in op else "")] += 1
    out = ["# SASS opcode evidence per kernel of `cdfo_b200/libcdfo_b200.so` (%s)" % rnd, "",
           "`cuobjdump -sass` of the shipped library, static instruction counts. `UTCHMMA` = `tcgen05.mma kind::f16`, `LDTM` / `STTM` = `tcgen05.ld` / "
           "`tcgen05.st`, `UTMALDG` / `UBLKCP` = tiled / bulk TMA loads, `UTCBAR` = `tcgen05.commit`, `TEX` = texture fetch, `HMMA` = warp-level "
           "`mma.sync` (legacy tensor path), `LDGSTS` = `cp.async`.", "", "| kernel | SASS instructions | tensor / TMA / TMEM opcodes |", "|---|---|---|"]
    for name, c in kernels.items():
        ops = ", ".join("%s x%d" % (k, v) for k, v in sorted(c.items()) if k != "_total" and not k.startswith(("MUFU", "SYNCS")))
        out.append("| `%s` | %d | %s |" % (name[:100], c["_total"], ops or "-"))
    path = os.path.join(ROOT, "profiles", "%s_sass_opcodes.md" % rnd)
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[6:]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
