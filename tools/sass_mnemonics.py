#!/usr/bin/env python
"""Static evidence, no GPU needed: per-kernel counts of the SASS mnemonics that show which hardware
paths the shipped libdronestep.so uses (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async, FFMA2 / FADD2 / FMUL2 =
packed f32, DFMA / DADD / DMUL = the fp64 chain, MUFU.RSQ64H / RCP64H = fp64 sqrt / division seeds).
Usage: python tools/sass_mnemonics.py [substring ...] > profiles/rNN/..._sass_mnemonics.txt"""
import collections, os, re, subprocess, sys

KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "DFMA", "DADD", "DMUL",
        "MUFU.RSQ64H", "MUFU.RCP64H", "BAR.SYNC", "STG.E.128", "STG.E.64", "LDS.128", "STS.128", "ATOMS", "REDUX",
        "SHFL", "VOTE", "LDG.E"]
DEFAULT = ("rollout_kernel<double, 2, 256, 1>", "rollout_kernel<float, 2, 256, 1>", "policy_kernel<double, 6>",
           "policy_kernel<double, 15>", "returns_kernel<double, 3>", "step_kernel<double, 2, 256>",
           "rollout_control_kernel<double, 2, 256>", "reset_random_kernel<double>", "reduce_agg_kernel")

def main(want):
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scalable_collision_avoidance_rl_b200", "libdronestep.so")
    sass = subprocess.run(["cuobjdump", "-sass", os.path.abspath(so)], capture_output=True, text=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for l in sass.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1); continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if cur and m:
            cnt[cur]["_total"] += 1
            for k in KEYS:
                if m.group(1).startswith(k): cnt[cur][k] += 1
    rows = []
    for f, c in cnt.items():
        name = re.sub(r"\(.*", "", subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip())
        if any(w in name for w in want):
            rows.append((name.replace("void ", ""), c))
    print(f"# SASS mnemonic counts per kernel of {os.path.basename(so)} (cuobjdump -sass, sm_100a); static, not executed counts")
    for name, c in sorted(rows):
        print(f"{name}: {c['_total']} instructions; " + ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]))

if __name__ == "__main__":
    main(tuple(sys.argv[1:]) or DEFAULT)
