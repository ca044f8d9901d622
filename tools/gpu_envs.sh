#!/bin/bash
OUT=gpurun_out/${1:-q02}; mkdir -p $OUT
for E in 1024 2048 3072 3552 4096 8192; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra --envs $E > $OUT/bench_$E.json 2>$OUT/bench_$E.err
python -c "import json;d=json.load(open('$OUT/bench_$E.json'));print($E, d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
done
