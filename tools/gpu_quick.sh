#!/bin/bash
OUT=gpurun_out/${1:-q01}; mkdir -p $OUT
DS_PLAN_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra ${@:2} > $OUT/bench.json 2>$OUT/bench.err
grep "resident" $OUT/bench.err | head -2
python -c "import json;d=json.load(open('$OUT/bench.json'));print(d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
