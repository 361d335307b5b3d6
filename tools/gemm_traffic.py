#!/usr/bin/env python3
"""DRAM bytes per launch of the tcgen05 layer GEMMs from an `ncu --set full` raw CSV (occurrence-weighted over the launches captured)
-> profiles/r02_gemm_traffic.json, which bench.py reads for `roofline.traffic`.
   python tools/gemm_traffic.py <raw.csv> <token rows> <source label> [leading GEMM launches to skip: the stem's]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, rows_key, label = sys.argv[1], sys.argv[2], sys.argv[3]
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ki = hdr.index("Kernel Name"); ri = hdr.index("dram__bytes_read.sum"); wi = hdr.index("dram__bytes_write.sum")
tot = n = 0
for r in rows[2:]:
    if "gemm_tc" not in r[ki] and "gemm_q8" not in r[ki]:
        continue
    if skip > 0:
        skip -= 1
        continue
    tot += float(r[ri].replace(",", "")) * UNIT.get(units[ri], 1.0) + float(r[wi].replace(",", "")) * UNIT.get(units[wi], 1.0); n += 1
out = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
d = json.load(open(out)) if os.path.exists(out) else {}
d[rows_key] = {"dram_bytes_per_launch": tot / max(n, 1), "launches": n, "source": label}
json.dump(d, open(out, "w"), indent=1)
print(d[rows_key])
