"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share table (markdown under profiles/)."""
import collections
import csv
import sys

OURS = ("cdfo::", "mdta::", "rs::", "dtex::", "pw::", "feat::", "cpair::", "c4::", "fdcn::", "fc8::", "lml::", "lct::")


def main(src, dst, title, note, marker=None, first=0, last=0):
    """marker / first / last: keep only the launches from the `first`-th to just before the `last`-th launch (1-based) of the kernel whose
    name contains `marker` -- e.g. the fused alignment kernel runs once per step, so (marker, 5, 7) = steps 5 and 6 of the process."""
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, n = collections.OrderedDict(), 0
    body = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
    if marker:
        hits = [i for i, r in enumerate(body) if marker in r[ik]]
        body = body[hits[first - 1]:hits[last - 1]]
    for r in body:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        a = agg.setdefault(r[ik], [0.0, 0])
        a[0] += v
        a[1] += 1
        n += 1
    tot = sum(a[0] for a in agg.values())
    ours = sum(t for k, (t, c) in agg.items() if any(o in k for o in OURS))
    out = ["# " + title, "", note, "", "| total us | launches | share | kernel |", "|---|---|---|---|"]
    for k, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:45]:
        out.append("| %.1f | %d | %.1f%% | `%s` |" % (t, c, 100 * t / tot, k[:110]))
    out += ["", "Total: %d launches, %.1f ms of kernel time; kernels of this repo (namespaces %s): %.1f%% of it."
            % (n, tot / 1e3, ", ".join("`%s`" % o for o in OURS), 100 * ours / tot)]
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[4:26]))
    print(out[-1])


if __name__ == "__main__":
    if len(sys.argv) > 5:
        main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], int(sys.argv[6]), int(sys.argv[7]))
    else:
        main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4])
