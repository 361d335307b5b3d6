#!/usr/bin/env python3
"""Top stall sites of an `ncu --page source --csv` dump (SASS view): per kernel, the stall-reason totals and the N instructions with
the most warp-stall samples.  usage: ncu_source_top.py file.csv [N]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) >= len(hdr) - 2:
                    body.append(rows[j])
                j += 1
            col = {h: k for k, h in enumerate(hdr)}
            stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

            def num(r, h):
                try:
                    return float(r[col[h]].replace(",", "")) if r[col[h]] else 0.0
                except (ValueError, IndexError):
                    return 0.0
            total = sum(num(r, "# Samples") for r in body)
            inst = sum(num(r, "Instructions Executed") for r in body)
            print(f"== {name}: {len(body)} SASS instructions, {total:.0f} samples, {inst:.0f} warp instructions executed")
            tot = {h: sum(num(r, h) for r in body) for h in stall_cols}
            print("   stalls: " + ", ".join(f"{h[6:]} {100 * v / max(total, 1):.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0.01 * total))
            body.sort(key=lambda r: -num(r, "# Samples"))
            for r in body[:top]:
                reasons = sorted(((num(r, h), h[6:]) for h in stall_cols), reverse=True)[:2]
                print(f"   {100 * num(r, '# Samples') / max(total, 1):5.1f}%  exec {num(r, 'Instructions Executed'):9.0f}  {r[col['Source']][:90]:90s} {reasons[0][1]} {reasons[1][1]}")
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main()
