#!/bin/bash
# scan one environment variable over values: tools/gpu_env_scan.sh TAG VAR v1 v2 ...
OUT=gpurun_out/$1; mkdir -p $OUT; VAR=$2; shift 2
for v in "$@"; do
env $VAR=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_$v.json 2>$OUT/bench_$v.err
python -c "import json;d=json.load(open('$OUT/bench_$v.json'));print('$VAR=$v',d['value'],d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],d['roofline']['frac'])"
done
