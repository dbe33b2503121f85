"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        val *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'total ms':>10} {'launches':>8} {'avg us':>9} {'share':>6}  kernel")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:10.3f} {n:8d} {1e3 * ms / n:9.1f} {100 * ms / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
