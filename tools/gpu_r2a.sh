#!/bin/bash
# round-2 first check: new bench under the driver's command line + launch-plan variants
OUT=gpurun_out/s01; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_driver.json 2> $OUT/bench_driver.err; echo "rc=$?"
cut -c1-400 $OUT/bench_driver.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"
cut -c1-300 $OUT/bench_ref.json
for plan in "1 16" "2 8" "4 4" "7 2" "8 2" "3 5" "1 8" "2 4"; do
  set -- $plan
  DS_PLAN_G=$1 DS_PLAN_TC=$2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/plan_$1_$2.json 2>$OUT/plan_$1_$2.err
  echo "G=$1 TC=$2: $(python -c "import json;d=json.load(open('$OUT/plan_$1_$2.json'));print(d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])")"
done
