"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
usage: python tools/summarize_ncu.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import collections
import sys


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        if unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        rows.append((r["Kernel Name"].split("(")[0], v))
    tot = sum(v for _, v in rows)
    agg = collections.OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f"launches profiled: {len(rows)}, total device time {tot/1e6:.3f} ms (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {v/1e3:.1f} | {100*v/tot:.1f}% | {v/1e3/n:.2f} |")


if __name__ == "__main__":
    main(sys.argv[1])
