#!/bin/bash
# kernel iteration with a scan of the segment-balance parameter: parity tests, config-3 bench per rho, ncu capture
OUT=gpurun_out/${1:-it}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rollout or dense or golden or config" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
for rho in ${RHOS:-70 50 35}; do
DS_RO2_RHO_PERMILLE=$rho DS_PLAN_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_rho$rho.json 2>$OUT/bench_rho$rho.err
python -c "import json;d=json.load(open('$OUT/bench_rho$rho.json'));print('rho$rho',d['value'],d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],d['roofline']['frac'])"
done
grep resident $OUT/bench_rho70.err | head -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout2_kernel -s 3 -c 1 \
    -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/ncu.log 2>&1
