#!/bin/bash
OUT=gpurun_out/${1:-s07}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for segs in 4 8 2; do
DS_RO2_SEGS=$segs DS_PLAN_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_s$segs.json 2>$OUT/bench_s$segs.err
grep resident $OUT/bench_s$segs.err | head -1
python -c "import json;d=json.load(open('$OUT/bench_s$segs.json'));print('segs $segs',d['value'],d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],d['roofline']['frac'])"
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra --log-mode 2 > $OUT/bench_lm2.json 2>$OUT/bench_lm2.err
python -c "import json;d=json.load(open('$OUT/bench_lm2.json'));print('lm2',d['value'],d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],d['roofline']['frac'])"
for w in config2 config4; do
DS_PLAN_DEBUG=1 timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_$w.json 2>$OUT/bench_$w.err
grep resident $OUT/bench_$w.err | head -1
python -c "import json;d=json.load(open('$OUT/bench_$w.json'));print('$w',d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout2_kernel -s 3 -c 1 \
    -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/ncu.log 2>&1
ls $OUT | wc -l
