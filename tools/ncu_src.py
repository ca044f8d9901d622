#!/usr/bin/env python
"""Per-source-line warp-instruction / stall-sample shares of one kernel in an .ncu-rep with the source
text beside them, in source order, plus totals per marked phase.  Read here, no GPU.
Usage: python tools/ncu_src.py REPORT.ncu-rep MANGLED_KERNEL_SUBSTRING [file-substring] [min_pct]"""
import csv, io, re, subprocess, sys, tempfile, os, collections
rep, kern = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else "rollout2"
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.15
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}; sass = rows[2:]
so = os.path.join(os.path.dirname(os.path.abspath(rep)), "lib.so")      # the library the report was taken with (tools/dev_cycle.sh)
if not os.path.exists(so):
    so = os.path.join(ROOT, "scalable_collision_avoidance_rl_b200", "libdronestep.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout
lines, cur, inside = [], None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        inside = kern in l; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip()); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l): lines.append((cur, l.strip()))
if len(lines) != len(sass): print(f"warning: {len(lines)} vs {len(sass)}", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
# attribute to the OUTERMOST line inside files matching fsub when the inline chain is available
for (loc, txt), r in zip(lines, sass):
    ie = int(r[col["Instructions Executed"]]); te = int(r[col["Thread Instructions Executed"]]); sm = int(r[col["# Samples"]])
    key = (loc[0], loc[1]) if loc else ("?", 0)
    a = agg[key]; a[0] += ie; a[1] += te; a[2] += sm
    tot[0] += ie; tot[1] += te; tot[2] += sm
print(f"total warp-instr {tot[0]}  thread-instr {tot[1]}  samples {tot[2]}")
files = collections.defaultdict(lambda: [0, 0, 0])
for (f, ln), a in agg.items():
    for q in range(3): files[f][q] += a[q]
for f, a in sorted(files.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:36s} instr% {100*a[0]/tot[0]:6.2f} samples% {100*a[2]/tot[2]:6.2f} thr {a[1]/max(a[0],1):5.1f}")
for f in sorted(files):
    if fsub not in f and "kernels" not in f: continue
    path = None
    for d in ("scalable_collision_avoidance_rl_b200/csrc",):
        p = os.path.join(ROOT, d, f)
        if os.path.exists(p): path = p
    src = open(path).read().splitlines() if path else []
    print(f"---- {f}")
    for (ff, ln), a in sorted(agg.items()):
        if ff != f: continue
        pct = 100 * a[0] / tot[0]; sp = 100 * a[2] / tot[2]
        if pct < minpct and sp < minpct: continue
        text = src[ln - 1].strip()[:90] if 0 < ln <= len(src) else ""
        print(f"{ln:5d} {pct:6.2f} {sp:6.2f} {a[1]/max(a[0],1):5.1f} | {text}")
