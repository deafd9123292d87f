"""Config c4 across GPU counts: the per-sequence PSNR / SSIM of two `bench.py --workload c4` lines must be IDENTICAL -- a sequence's
result may not depend on which rank ran it or on what shared its batch (Gumbel noise is keyed by (sequence, frame, neighbour), every
kernel reduces in a fixed order per sample).  Usage: python tools/compare_c4.py a.json b.json [more.json ...]; exit code 1 on a mismatch."""
import json
import sys


def load(path):
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith("{"):
                return json.loads(line)
    raise SystemExit("%s: no JSON line" % path)


def main(paths):
    lines = [load(p) for p in paths]
    ref = lines[0]
    bad = 0
    for p, d in zip(paths, lines):
        n = len(d["psnr"])
        dp = max(abs(a - b) for a, b in zip(d["psnr"], ref["psnr"]))
        ds = max(abs(a - b) for a, b in zip(d["ssim"], ref["ssim"]))
        same = d["psnr"] == ref["psnr"] and d["ssim"] == ref["ssim"] and n == len(ref["psnr"])
        bad += not same
        print("%-40s n_gpus %d  %7.1f fps  %d sequences  max|dPSNR| %.3g dB  max|dSSIM| %.3g  %s"
              % (p, d["n_gpus"], d["value"], n, dp, ds, "identical" if same else "DIFFERENT"))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
