#!/bin/bash
# config-5 (n = 128) plan scan of the general rollout kernel + one full ncu capture
OUT=gpurun_out/${1:-c5}; mkdir -p $OUT
run() { tag=$1; shift; env "$@" DS_PLAN_DEBUG=1 timeout 300 python bench.py --workload config5 --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/b_$tag.json 2>$OUT/b_$tag.err
python -c "import json;d=json.load(open('$OUT/b_$tag.json'));print('$tag',d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"; grep "rollout plan" $OUT/b_$tag.err | head -1; }
run base A=1
run tc2 DS_PLAN_G=1 DS_PLAN_TC=2
run lpr3 DS_PLAN_LPR=3
run lpr2 DS_PLAN_LPR=2
run inl0 DS_PLAN_INLINE=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o $OUT/prof python bench.py --workload config5 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/ncu.log 2>&1; tail -1 $OUT/ncu.log
