"""`ncu --page raw --csv` dump -> per-kernel summary (time, DRAM bytes, achieved GB/s, pipe utilisation) and profiles/ncu_traffic.json.

    python tools/ncu_traffic.py gpurun_out/r02_layer_raw.csv <name of the copy kept under profiles/>
The LAST launch of each distinct (kernel name, grid) is reported (tools/gpu_layer_ncu.py launches everything twice: the first is a
warm-up)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, kept = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1]))
rows = list(csv.reader(open(path)))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}


def f(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return None


last = {}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]]
    key = (name, r[col["Grid Size"]] if "Grid Size" in col else "")
    last[key] = r
out = []
for (name, grid), r in last.items():
    t = f(r, "gpu__time_duration.sum")
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    unit_t = rows[1][col["gpu__time_duration.sum"]]
    unit_b = rows[1][col["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in col else "byte"
    scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit_t, 1.0)
    scale_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_b, 1)
    us = None if t is None else t * scale_t
    by = None if rd is None else (rd + wr) * scale_b
    short = re.sub(r"\(.*", "", name)
    out.append({"kernel": short, "grid": grid, "us": us, "dram_bytes": by,
                "dram_gbs": None if not us or by is None else by / us / 1e3,
                "tensor_pipe_pct": f(r, "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sm_active") or f(r, "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sm_active"),
                "sm_throughput_pct": f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                "dram_throughput_pct": f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "achieved_occupancy_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "regs": f(r, "launch__registers_per_thread")})
for o in out:
    print(f"{o['kernel'][:70]:70s} grid {o['grid']:>14s} {o['us'] or 0:9.1f} us  dram {0 if o['dram_bytes'] is None else o['dram_bytes'] / 1e6:9.1f} MB "
          f"{o['dram_gbs'] or 0:7.0f} GB/s  sm {o['sm_throughput_pct']} dram% {o['dram_throughput_pct']} occ {o['achieved_occupancy_pct']}")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ncu_kernel_summary.json"), "w"), indent=1)
