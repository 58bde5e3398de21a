"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    m = re.search(r"(k_[a-z0-9_]+)", name)
    if m:
        return m.group(1)
    if "at::" in name or "elementwise" in name:
        return "torch_elementwise/copy/fill"
    return name[:60]


def main(path: str) -> None:
    tot = defaultdict(lambda: [0, 0.0])
    with open(path, newline="") as fh:
        rows = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(rows)
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        k = short(r["Kernel Name"])
        tot[k][0] += 1
        tot[k][1] += ns
        n += 1
    total = sum(v[1] for v in tot.values()) or 1.0
    out = {"launches": n, "total_ms": total / 1e6,
           "kernels": {k: {"launches": v[0], "ms": v[1] / 1e6, "share": v[1] / total, "avg_us": v[1] / v[0] / 1e3}
                       for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
