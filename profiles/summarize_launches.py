"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list into
per-kernel totals, shares and (when the DRAM metrics are present) DRAM bytes per launch.

  python profiles/summarize_launches.py gpurun_out/launches.csv profiles/round1_launches.md [profiles/round1_traffic.json]
"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, out, traffic_out=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, mi, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    agg = collections.OrderedDict()      # kernel -> {launch ids, time_us, dram bytes}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"^void ", "", r[ki]).replace("<unnamed>::", "")
        name = re.sub(r"\(.*", "", name)                 # drop the argument list, keep template arguments
        try:
            v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        except ValueError:
            continue
        d = agg.setdefault(name, dict(ids=set(), us=0.0, dram=0.0))
        d["ids"].add(r[ii])
        if r[mi].startswith("gpu__time_duration"):
            d["us"] += v
        elif r[mi].startswith("dram__bytes"):
            d["dram"] += v
    tot = sum(d["us"] for d in agg.values())
    has_dram = any(d["dram"] for d in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary: {path}\n\n{sum(len(d['ids']) for d in agg.values())} launches, {tot / 1e3:.2f} ms total "
                "(cold-cache, serialised: compare SHARES, not absolutes)\n\n| kernel | launches | total ms | share | avg us |"
                + (" DRAM MB / launch |" if has_dram else "") + "\n|---|---:|---:|---:|---:|" + ("---:|" if has_dram else "") + "\n")
        for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            n = len(d["ids"])
            f.write(f"| `{k[:90]}` | {n} | {d['us'] / 1e3:.3f} | {100 * d['us'] / tot:.1f}% | {d['us'] / n:.1f} |"
                    + (f" {d['dram'] / n / 1e6:.2f} |" if has_dram else "") + "\n")
    if traffic_out and has_dram:
        json.dump({k: dict(launches=len(d["ids"]), dram_bytes_per_launch=d["dram"] / len(d["ids"]), avg_us=d["us"] / len(d["ids"]))
                   for k, d in agg.items()}, open(traffic_out, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    main(*sys.argv[1:4])
