#!/bin/bash
# config-2 (n = 5) scan of the segments per environment
OUT=gpurun_out/${1:-c2}; mkdir -p $OUT
for sg in 0 2 1 8; do
E=""; [ $sg != 0 ] && E="DS_RO2_SEGS=$sg"
env $E DS_PLAN_DEBUG=1 timeout 300 python bench.py --workload config2 --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/b_s$sg.json 2>$OUT/b_s$sg.err
python -c "import json;d=json.load(open('$OUT/b_s$sg.json'));print('config2 segs$sg',d['value'],d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],d['roofline']['frac'])"
grep "resident" $OUT/b_s$sg.err | head -1
done
