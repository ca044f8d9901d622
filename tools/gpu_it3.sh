#!/bin/bash
# kernel iteration with a scan of the segments per environment: parity tests, config-3 / config-4 bench per value, ncu capture
OUT=gpurun_out/${1:-it}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rollout or dense or golden or config" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
for w in config3 config4; do
for sg in ${SEGS:-0 1 2 4}; do
E=""; [ $sg != 0 ] && E="DS_RO2_SEGS=$sg"
env $E DS_PLAN_DEBUG=1 timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_${w}_s$sg.json 2>$OUT/bench_${w}_s$sg.err
python -c "import json;d=json.load(open('$OUT/bench_${w}_s$sg.json'));print('$w segs$sg',d['value'],d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],d['roofline']['frac'])"
grep "plan:\|resident" $OUT/bench_${w}_s$sg.err | head -2
done; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout2_kernel -s 3 -c 1 \
    -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/ncu.log 2>&1
