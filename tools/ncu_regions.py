#!/usr/bin/env python
"""Per-code-region instruction / stall-sample shares of the rollout kernel in an .ncu-rep (read
without a GPU): the report's SASS page joined with `nvdisasm --print-line-info` of the cubin inside
the shipped .so, aggregated by the source function / kernel phase each instruction's innermost
line belongs to.  Usage: python tools/ncu_regions.py REPORT.ncu-rep MANGLED_KERNEL_SUBSTRING [lib.so]"""
import csv, io, re, subprocess, sys, collections, os, tempfile
rep, kern = sys.argv[1], sys.argv[2]
src = open('/root/repo/scalable_collision_avoidance_rl_b200/csrc/dronestep_kernels.cuh').read().splitlines()
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr=rows[1]; col={h:i for i,h in enumerate(hdr)}; sass=rows[2:]
so=sys.argv[3] if len(sys.argv)>3 else '/root/repo/scalable_collision_avoidance_rl_b200/libdronestep.so'
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump","-xelf","all",so],cwd=td,capture_output=True)
    cub=[f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis=subprocess.run(["nvdisasm","--print-line-info",os.path.join(td,cub)],capture_output=True,text=True).stdout
lines=[];cur=None;inside=False
for l in dis.splitlines():
    if l.startswith(".text."): inside = kern in l; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur=(os.path.basename(m.group(1)), int(m.group(2)), m.group(3)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l): lines.append((cur,l.strip()))
assert len(lines)==len(sass),(len(lines),len(sass))
# function ranges by scanning source for names
def find(pat):
    for i,l in enumerate(src):
        if pat in l: return i+1
    return None
marks=[(1,'helpers'),(find('DS_HD double log_r'),'log_r'),(find('inline void fill_log_table'),'x'),(find('DS_HD void eval_pair'),'eval_pair'),(find('DS_HD unsigned pack_entry'),'pack'),(find('DS_HD bool topk_offer'),'topk_offer'),(find('DS_HD void row_begin'),'row_begin'),(find('DS_HD bool row_fold'),'row_fold'),(find('DS_HD void row_end'),'row_end'),(find('DS_HD void eval_row('),'eval_row'),(find('DS_HD void eval_row_from_list'),'eval_row_from_list'),(find('DS_HD void topk_insert'),'topk_insert'),(find('DS_HD void eval_row_near32'),'eval_row_near32'),(find('pass1_block(const'),'pass1_block'),(find('pass1_block_f32x2(const'),'pass1_f32x2'),(find('void cp_async_action(double2'),'cp_async'),(find('DS_HD int executed_slices'),'executed_slices'),(find('DS_HD void write_obs'),'write_obs'),(find('struct CtaSmem'),'step'),(find('rollout_kernel(const RolloutArgs'),'ro:prologue'),(find('// (a)'),'ro:(a)'),(find('// (b) sequential'),'ro:(b)'),(find('// (c) pass 1'),'ro:(c) pass1'),(find('// segment of the work list'),'ro:(c) scan+write'),(find('// (d) pass 2'),'ro:(d)'),(find('// (e) rows'),'ro:(e)'),(find('// (f) frame leaders'),'ro:(f)'),(find('// (g) stores'),'ro:(g)'),(find('// episode sums of this chunk'),'ro:acc'),(find('// Deterministic sum'),'end')]
marks=[m for m in marks if m[0]]
marks.sort()
def region(f,ln):
    if not f.endswith('dronestep_kernels.cuh'): return 'other:'+f
    r='?'
    for a,name in marks:
        if ln>=a: r=name
    return r
agg=collections.defaultdict(lambda:[0,0,0]); tot=[0,0,0]
for (loc,txt),r in zip(lines,sass):
    ie=int(r[col["Instructions Executed"]]); te=int(r[col["Thread Instructions Executed"]]); smp=int(r[col["# Samples"]])
    # inlined-at chain: use outermost? use innermost function
    key=region(loc[0],loc[1]) if loc else 'noline'
    a=agg[key]; a[0]+=ie;a[1]+=te;a[2]+=smp; tot[0]+=ie;tot[1]+=te;tot[2]+=smp
print("total warp-instr",tot[0],"samples",tot[2])
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:28s} instr% {100*a[0]/tot[0]:6.2f}  samples% {100*a[2]/tot[2]:6.2f}  thr {a[1]/max(a[0],1):5.1f}")
