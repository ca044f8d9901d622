#!/bin/bash
# shortest kernel iteration: a few parity tests + config-3 bench (no ncu)
OUT=gpurun_out/${1:-it}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_instantiation or golden or returns_recipe" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $OUT/pytest_gpu.log
for rho in ${RHOS:-55}; do
DS_RO2_RHO_PERMILLE=$rho timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/bench_rho$rho.json 2>$OUT/bench_rho$rho.err
python -c "import json;d=json.load(open('$OUT/bench_rho$rho.json'));print('rho$rho',d['value'],d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],d['roofline']['frac'])"
done
