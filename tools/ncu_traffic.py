"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into profiles/r2_ncu_traffic.json: per kernel the DRAM bytes
per launch (dram__bytes_read.sum + dram__bytes_write.sum), duration, tensor-pipe and DRAM utilisation.
usage: python tools/ncu_traffic.py name=report.ncu-rep[:kernel-regex] ..."""
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "sm__inst_executed_pipe_tensor.sum": "tensor_inst", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "launch__registers_per_thread": "regs", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "lts__t_bytes.sum": "l2_bytes"}


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
            "second": 1.0}.get(u, 1)


def summarise(rep, pattern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pattern and not re.search(pattern, name):
            continue
        e = {"kernel": name[:120], "grid": r[col.get("Grid Size", 0)]}
        for m, k in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                e[k] = float(r[col[m]].replace(",", "")) * unit_scale(units[col[m]])
        res.append(e)
    return res


if __name__ == "__main__":
    p = ROOT / "profiles" / "r2_ncu_traffic.json"
    data = json.loads(p.read_text()) if p.exists() else {}
    for arg in sys.argv[1:]:
        name, rest = arg.split("=", 1)
        rep, _, pat = rest.partition(":")
        rows = summarise(rep, pat)
        if not rows:
            print("no kernels matched for", name)
            continue
        n = len(rows)
        agg = {"report": Path(rep).name, "launches": n, "kernel": rows[0]["kernel"], "grid": rows[0]["grid"]}
        for k in ("dram_read", "dram_write", "duration", "tensor_pipe_pct", "dram_pct", "sm_pct", "l2_bytes", "regs"):
            vals = [r[k] for r in rows if k in r]
            if vals:
                agg[k + "_avg"] = sum(vals) / len(vals)
        agg["dram_bytes_per_launch"] = agg.get("dram_read_avg", 0) + agg.get("dram_write_avg", 0)
        data[name] = agg
        print(name, json.dumps(agg))
    p.write_text(json.dumps(data, indent=1))
