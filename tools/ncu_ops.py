#!/usr/bin/env python
"""Executed warp-instructions of one kernel in an .ncu-rep grouped by SASS opcode. Usage: ncu_ops.py REP [kernel-substr]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}; sass = rows[2:]
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
for r in sass:
    src = r[col["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:12]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "ATOM", "RED", "MUFU", "SHFL", "BAR", "SYNCS", "UBLKCP")) else op.split(".")[0]
    ie = int(r[col["Instructions Executed"]]); te = int(r[col["Thread Instructions Executed"]]); sm = int(r[col["# Samples"]])
    a = agg[op]; a[0] += ie; a[1] += te; a[2] += sm
    tot[0] += ie; tot[1] += te; tot[2] += sm
print(f"total warp-instr {tot[0]} samples {tot[2]}")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"{op:16s} instr% {100*a[0]/tot[0]:6.2f} samples% {100*a[2]/max(tot[2],1):6.2f} thr {a[1]/max(a[0],1):5.1f}")
