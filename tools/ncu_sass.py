#!/usr/bin/env python
"""SASS of one kernel in an .ncu-rep in program order with executed counts (per `unit` launches of the
loop body), active threads and stall samples. Usage: ncu_sass.py REP unit_count [min_ratio]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; unit = float(sys.argv[2]); minr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}; sass = rows[2:]
acc = 0.0
for k, r in enumerate(sass):
    ie = int(r[col["Instructions Executed"]]); te = int(r[col["Thread Instructions Executed"]]); sm = int(r[col["# Samples"]])
    ratio = ie / unit
    acc += ratio
    if ratio >= minr:
        print(f"{k:5d} {ratio:6.2f} {te/max(ie,1):5.1f} {sm:5d} {acc:8.1f}  {r[col['Source']][:100]}")
