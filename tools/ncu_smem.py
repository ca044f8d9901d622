#!/usr/bin/env python
"""Shared-memory wavefronts (actual / ideal) per source line of one kernel in an .ncu-rep, joined with
the SASS of the library the report was taken with (lib.so next to the report).  Read here, no GPU.
Usage: python tools/ncu_smem.py REPORT.ncu-rep MANGLED_KERNEL_SUBSTRING [min_pct]"""
import csv, io, re, subprocess, sys, tempfile, os, collections
rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}; sass = rows[2:]
so = os.path.join(os.path.dirname(os.path.abspath(rep)), "lib.so")
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
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l): lines.append((cur, l.strip()))
src = {}
def text(f, ln):
    if f not in src:
        p = os.path.join(ROOT, "scalable_collision_avoidance_rl_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
agg = collections.defaultdict(lambda: [0, 0, 0, set()]); tot = [0, 0]
for (loc, txt), r in zip(lines, sass):
    w = int(r[col["L1 Wavefronts Shared"]] or 0); wi = int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    if not w: continue
    a = agg[loc]; a[0] += w; a[1] += wi; a[2] += int(r[col["Instructions Executed"]]); a[3].add(txt.split()[1] if len(txt.split()) > 1 else "")
    tot[0] += w; tot[1] += wi
print(f"shared wavefronts {tot[0]}  ideal {tot[1]}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    pct = 100.0 * a[0] / tot[0]
    if pct < minpct: continue
    print(f"{loc[0][:22]:22s} {loc[1]:5d} {pct:6.2f}%  wf/instr {a[0] / max(a[2], 1):5.2f}  ideal {a[1] / max(a[2], 1):5.2f}  {','.join(sorted(a[3]))[:30]:30s} | {text(*loc)}")
