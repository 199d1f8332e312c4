"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import collections
import csv
import re
import sys


def main(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"^void ", "", r[ki])
        name = re.sub(r"\(.*", "", name).replace("<unnamed>::", "")
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(d[1] for d in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary: {path}\n\n{sum(d[0] for d in agg.values())} launches, {tot / 1e3:.2f} ms total "
                "(cold-cache, serialised: compare SHARES, not absolutes)\n\n| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {n} | {t / 1e3:.3f} | {100 * t / tot:.1f}% | {t / n:.1f} |\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
