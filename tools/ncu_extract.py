#!/usr/bin/env python
"""Key metrics of every kernel in .ncu-rep files (`ncu --set full`), as a markdown table / JSON.

    python tools/ncu_extract.py gpurun_out/r02_*.ncu-rep [--json]
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "sm__cycles_active.avg": "sm_cycles",
    "smsp__cycles_active.avg": "smsp_cycles",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "smsp__issue_active.avg.pct": "issue_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_hz",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_bytes.sum": "l2_bytes",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "hz": 1.0, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9, "cycle/nsecond": 1e9,
        "cycle/usecond": 1e6, "cycle/second": 1.0}


def short(name):
    m = re.search(r"((?:c3::)?[A-Za-z_0-9]+_kernel(?:<[^>]*>)?)", name)
    return m.group(1) if m else name[:50]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    res = []
    for r in body:
        d = {"file": path.split("/")[-1], "kernel": short(r[hdr.index("Kernel Name")])}
        for k, v in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[v] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                except ValueError:
                    pass
        res.append(d)
    return res


def main():
    files = [a for a in sys.argv[1:] if not a.startswith("--")]
    allr = [x for f in files for x in load(f)]
    if "--json" in sys.argv:
        print(json.dumps(allr, indent=1))
        return
    print("| file | kernel | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % | tensor % | L2 % | L2 hit % | issue % | occ % | SM GHz | regs | grid x block |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for d in allr:
        t = d.get("time_us", 0.0)
        rd, wr = d.get("dram_rd", 0.0), d.get("dram_wr", 0.0)
        gbs = (rd + wr) / t / 1e3 if t else 0.0
        print(f"| {d['file'].replace('.ncu-rep', '')} | `{d['kernel']}` | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | "
              f"{d.get('dram_pct', 0):.1f} | {d.get('tensor_pct', 0):.1f} | {d.get('l2_pct', 0):.1f} | {d.get('l2_hit_pct', 0):.1f} | "
              f"{d.get('issue_pct', 0):.1f} | {d.get('occ_pct', 0):.1f} | {d.get('sm_hz', 0) / 1e9:.2f} | {int(d.get('regs', 0))} | "
              f"{int(d.get('grid', 0))} x {int(d.get('block', 0))} |")


if __name__ == "__main__":
    main()
