#!/bin/bash
# rollout2 iteration: quick parity subset, bench, ncu capture
OUT=gpurun_out/${1:-s03}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for mode in 0 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra --log-mode $mode > $OUT/bench_lm$mode.json 2>$OUT/bench_lm$mode.err; 
python -c "import json;d=json.load(open('$OUT/bench_lm$mode.json'));print('lm$mode',d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
done
for w in config2 config4; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_$w.json 2>$OUT/bench_$w.err
python -c "import json;d=json.load(open('$OUT/bench_$w.json'));print('$w',d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout2_kernel -s 3 -c 1 \
    -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/ncu.log 2>&1
ls -la $OUT | head -20
