#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a
markdown table (per-kernel count / total / share).  Usage:
    python tools/summarize_launches.py gpurun_out/launches.csv [title] > profiles/<name>.md
ncu times are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    print(f"# {title}\n")
    print(f"{n} launches, {tot / 1e3:.3f} ms summed device time (ncu, serialised, cold cache, --clock-control none)\n")
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:90]}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    main()
