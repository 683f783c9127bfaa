"""Per (kernel, grid) averages of an `ncu --metrics ... --csv` log (long format: one row per launch and metric).
usage: python tools/ncu_csv_by_kernel.py file.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
c = {h: i for i, h in enumerate(hdr)}
acc = defaultdict(lambda: defaultdict(list))
for r in rows[1:]:
    if r[c["ID"]] == "ID":
        continue
    key = (r[c["Kernel Name"]][:70], r[c["Grid Size"]])
    try:
        acc[key][(r[c["Metric Name"]], r[c["Metric Unit"]])].append(float(r[c["Metric Value"]].replace(",", "")))
    except ValueError:
        pass
for key, ms in sorted(acc.items(), key=lambda kv: -len(next(iter(kv[1].values())))):
    n = len(next(iter(ms.values())))
    print(f"{key[0]} grid={key[1]}  launches={n}")
    for (m, u), vals in ms.items():
        print(f"    {m:75s} {sum(vals) / len(vals):14.3f} {u}")
