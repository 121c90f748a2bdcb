"""Aggregates an ncu report's source page per CUDA source line (stall samples, instructions).
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, agg = None, None, []
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not r or hdr is None or r[0] == "Function Name":
        continue
    if r[0].strip().isdigit() and r[2] == "-":
        try:
            samp = int(r[4])
            inst = int(r[hdr.index("Instructions Executed")])
        except Exception:
            continue
        agg.append((samp, inst, cur, int(r[0]), r[1].strip()[:100]))
tot = sum(a[0] for a in agg)
print("total samples", tot, "instructions", sum(a[1] for a in agg))
for a in sorted(agg, reverse=True)[:top]:
    print("%6d %5.1f%% %9d  %s:%d  %s" % (a[0], 100.0 * a[0] / max(tot, 1), a[1], a[2], a[3], a[4]))
