"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (cold-cache, serialised
per-launch times: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*$", "", row["Kernel Name"])
        t = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        t = t / 1000.0 if unit == "ns" else t * 1000.0 if unit == "ms" else t
        key = name + " grid=" + row["Grid Size"].replace(" ", "") if "--grid" in sys.argv else name
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot/1000:.3f} ms summed kernel time")
    print(f"# {'total_us':>10} {'launches':>8} {'avg_us':>9} {'share':>6}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{a[1]:12.1f} {a[0]:8d} {a[1]/a[0]:9.2f} {100*a[1]/tot:5.1f}%  {k[:120]}")


if __name__ == "__main__":
    main(sys.argv[1])
