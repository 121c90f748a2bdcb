#!/usr/bin/env python
"""Aggregates an ncu report's source page by CUDA source line: stall samples and executed
instructions per line (needs -lineinfo and --import-source on).  usage: ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur, hdr = None, None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8 or r[2] != "-":
            continue
        try:
            line, samp, inst = int(r[0]), int(r[6] or 0), int(r[7] or 0)
        except ValueError:
            continue
        a = agg.setdefault((cur, line), [0, 0, r[1][:100]])
        a[0] += samp
        a[1] += inst
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print("total samples", tot, "instructions", toti)
    for (f, l), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%s:%4d samp %6d (%4.1f%%) inst %9d (%4.1f%%) | %s" % (f, l, s, 100.0 * s / tot, i, 100.0 * i / toti, src))


if __name__ == "__main__":
    main()
