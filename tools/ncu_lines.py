#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples of one kernel in an .ncu-rep, read without a GPU.
The report's SASS page (ncu --page source --csv) is joined with `nvdisasm --print-line-info` of the
cubin in the shipped .so (instruction order is identical), then aggregated by source line.
Usage: python tools/ncu_lines.py REPORT.ncu-rep MANGLED_KERNEL_SUBSTRING [top]"""
import csv, io, re, subprocess, sys, tempfile, os, collections

def main(rep, kern, top=45):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
    sass = rows[2:]
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scalable_collision_avoidance_rl_b200", "libdronestep.so")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout
    lines, cur, inside = [], None, False
    for l in dis.splitlines():
        if l.startswith(".text."):
            inside = kern in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip())
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
            lines.append((cur, l.strip()))
    if len(lines) != len(sass):
        print(f"warning: {len(lines)} disassembled instr vs {len(sass)} in report", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for (loc, txt), r in zip(lines, sass):
        ie = int(r[col["Instructions Executed"]]); te = int(r[col["Thread Instructions Executed"]]); sm = int(r[col["# Samples"]])
        key = (loc[0], loc[1]) if loc else ("?", 0)
        a = agg[key]; a[0] += ie; a[1] += te; a[2] += sm
        tot[0] += ie; tot[1] += te; tot[2] += sm
    print(f"total warp-instr {tot[0]}  thread-instr {tot[1]}  samples {tot[2]}")
    print(f"{'file:line':40s} {'warp-instr%':>11s} {'samples%':>9s} {'avg thr':>8s}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"{key[0]+':'+str(key[1]):40s} {100*a[0]/tot[0]:11.2f} {100*a[2]/max(tot[2],1):9.2f} {a[1]/max(a[0],1):8.1f}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 45)
