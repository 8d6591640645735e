#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv --log-file X.csv python bench.py ...`) per kernel instantiation.

    python tools/summarize_launches.py profiles/r02_launches.csv [--md] > profiles/r02_launches_summary.json

The JSON is what bench.py reads for `roofline.traffic` (mean DRAM bytes per launch of the dominant kernel, from the
CURRENT build's committed launch list, never a literal).  Times under ncu are cold-cache and serialised: compare SHARES.
"""
import csv
import json
import re
import sys
from collections import defaultdict


def short_name(full: str) -> str:
    m = re.search(r"(?:mscnn::)?((?:c3::)?[A-Za-z_0-9]+_kernel(?:<[^>]*>)?)", full)
    if m and ("mscnn" in full or "c3::" in full):
        n = m.group(1).replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")
        return re.sub(r"(conv_igemm_kernel<\d+), ([01])>", lambda k: f"{k.group(1)}, {'true' if k.group(2) == '1' else 'false'}>", n)
    return "(other: torch fills / one-off)"


def main():
    path = sys.argv[1]
    rows = defaultdict(dict)
    names = {}
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
                 "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        rows[r["ID"]][r["Metric Name"]] = v * scale
        names[r["ID"]] = short_name(r["Kernel Name"])
    agg = defaultdict(lambda: {"launches": 0, "time_us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    for i, m in rows.items():
        a = agg[names[i]]
        a["launches"] += 1
        a["time_us"] += m.get("gpu__time_duration.sum", 0.0)
        a["dram_read"] += m.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += m.get("dram__bytes_write.sum", 0.0)
    total = sum(a["time_us"] for a in agg.values()) or 1.0
    out = {"source": path, "total_ms": total / 1e3, "kernels": {}}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_us"]):
        n = a["launches"]
        out["kernels"][k] = {"launches": n, "total_ms": round(a["time_us"] / 1e3, 3), "share": round(a["time_us"] / total, 4),
                             "mean_us": round(a["time_us"] / n, 2),
                             "dram_read_bytes_per_launch": round(a["dram_read"] / n),
                             "dram_write_bytes_per_launch": round(a["dram_write"] / n),
                             "dram_gbs": round((a["dram_read"] + a["dram_write"]) / max(a["time_us"], 1e-9) / 1e3, 1)}
    if "--md" in sys.argv:
        print("| kernel | launches | total ms | share | mean us | DRAM rd MB/launch | DRAM wr MB/launch | DRAM GB/s |")
        print("|---|---:|---:|---:|---:|---:|---:|---:|")
        for k, v in out["kernels"].items():
            print(f"| `{k}` | {v['launches']} | {v['total_ms']} | {100 * v['share']:.1f}% | {v['mean_us']} | "
                  f"{v['dram_read_bytes_per_launch'] / 1e6:.1f} | {v['dram_write_bytes_per_launch'] / 1e6:.1f} | {v['dram_gbs']} |")
    else:
        print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
