#!/bin/bash
# rollout2 check: parity tests, then bench with the new / old kernel
OUT=gpurun_out/${1:-s02}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
for mode in 0 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra --log-mode $mode > $OUT/bench_new_lm$mode.json 2>$OUT/bench_new_lm$mode.err; echo "rc=$?"
python -c "import json;d=json.load(open('$OUT/bench_new_lm$mode.json'));print('new lm$mode',d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
done
DS_RO2=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_old.json 2>$OUT/bench_old.err
python -c "import json;d=json.load(open('$OUT/bench_old.json'));print('old',d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
DS_RO2_BULK=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_nobulk.json 2>$OUT/bench_nobulk.err
python -c "import json;d=json.load(open('$OUT/bench_nobulk.json'));print('nobulk',d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
for w in config2 config4; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_$w.json 2>$OUT/bench_$w.err
python -c "import json;d=json.load(open('$OUT/bench_$w.json'));print('$w',d['value'],d['roofline']['median_launch_ms'],d['roofline']['frac'])"
done
