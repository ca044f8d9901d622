#!/usr/bin/env python
"""Static SASS instruction counts per source line of one kernel in a cubin (nvdisasm --print-line-info),
attributed to the OUTERMOST inlined-at line inside `file-substring` when available.
Usage: sass_lines.py file.cubin kernel-substr [file-substring]"""
import re, subprocess, sys, collections, os
cub, kern = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else "rollout2"
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cub], capture_output=True, text=True).stdout
if not dis:
    dis = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
inside = False; cur = None; chain = []
cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter); total = 0
for l in dis.splitlines():
    if l.startswith(".text."):
        inside = kern in l; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f, ln, rest = os.path.basename(m.group(1)), int(m.group(2)), m.group(3)
        if "inlined at" in rest or not chain or True:
            pass
        if "inlined at" in l and chain:
            chain.append((f, ln))
        else:
            chain = [(f, ln)]
        # nvdisasm prints innermost first then 'inlined at' lines following; keep the last one in fsub
        continue
    m2 = re.search(r'//## .*inlined at "([^"]+)", line (\d+)', l)
    if m2:
        chain.append((os.path.basename(m2.group(1)), int(m2.group(2)))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        key = None
        for f, ln in reversed(chain):
            if fsub in f: key = (f, ln); break
        if key is None: key = chain[0] if chain else ("?", 0)
        cnt[key] += 1; total += 1
        op = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        ops[key][op.group(1).split(".")[0] if op else "?"] += 1
print("total static instructions", total)
src = {}
for (f, ln), c in sorted(cnt.items()):
    if f not in src:
        p = os.path.join("/root/repo/scalable_collision_avoidance_rl_b200/csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ""
    top = " ".join(f"{o}:{k}" for o, k in ops[(f, ln)].most_common(4))
    print(f"{f[-14:]}:{ln:4d} {c:4d}  {text:80s} {top}")
