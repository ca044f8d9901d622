#!/bin/bash
# the driver's command lines: both arms, N = 1
OUT=gpurun_out/${1:-b01}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "ours rc=$?"
python - <<PY
import json
d=json.load(open('$OUT/bench.json')); r=json.load(open('$OUT/bench_reference.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'],'launch ms',d['roofline']['median_launch_ms'])
print('e2e',d['e2e']['value'],d['e2e']['pcie_gbs_this_rank'],'full',d['e2e_full_precision_obs']['value'])
print('cpu port',d['cpu_baseline']['value'],'ref',d['cpu_baseline_reference']['value'],d['cpu_baseline_reference']['one_process']['value'])
print('reference arm',r['value'],r['cpu_baseline']['kind'])
for k,v in d['extra'].items(): print(k,v['value'],v['roofline']['frac'],v['roofline']['median_launch_ms'])
print('agg',d['agg_check'],'launches',d['gpu_launches'],d['clocks'])
PY
