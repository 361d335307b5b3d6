#!/usr/bin/env python3
"""Summarise an `ncu --set full` report: one line per profiled launch with the counters the roofline argument needs.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x_summary.txt]
Peaks: MEASURED_PEAKS.json (HBM GB/s, bf16 TFLOP/s) when present."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = {
    "dur_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pct_el": "sm__inst_executed_pipe_tensor.sum",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "smem_dyn": "launch__shared_mem_per_block_dynamic",
    "smem_sta": "launch__shared_mem_per_block_static",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    print(f"# {os.path.basename(rep)}  (HBM peak used: {hbm:.0f} GB/s {'measured' if peaks else 'fallback'}; per-launch, cold-cache, serialised by ncu)")
    print(f"{'kernel':58s} {'grid':>14s} {'us':>8s} {'dramMB':>8s} {'GB/s':>7s} {'%hbm':>6s} {'tensor%':>8s} {'l2%':>6s} {'sm%':>6s} {'warps%':>7s} {'regs':>5s} {'smemKB':>7s}")
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    for r in rows[2:]:
        def val(k):
            if k not in idx or r[idx[k]] in ("", "n/a"):
                return None
            v = float(r[idx[k]].replace(",", ""))
            u = units[idx[k]].split("/")[0]
            return v * UNIT.get(u, 1.0)
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("nsb::", "").replace("<unnamed>::", "")
        dur = val("dur_us")
        mb = ((val("dram_rd") or 0) + (val("dram_wr") or 0)) / 1e6
        gbs = mb / 1e3 / (dur * 1e-6) if dur else 0.0
        smem = ((val("smem_dyn") or 0) + (val("smem_sta") or 0)) / 1e3
        print(f"{name[:58]:58s} {r[gi]:>14s} {dur:8.2f} {mb:8.2f} {gbs:7.0f} {100 * gbs / hbm:6.1f} {val('tensor_pct') or 0:8.1f} {val('l2_pct') or 0:6.1f} "
              f"{val('sm_pct') or 0:6.1f} {val('warps_pct') or 0:7.1f} {int(val('regs') or 0):5d} {smem:7.1f}")


if __name__ == "__main__":
    main()
