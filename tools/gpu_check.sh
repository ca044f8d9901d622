#!/bin/bash
# GPU parity tests, smoke, bench (ours + reference arm), the other BASELINE configs, ncu launch list
# (stage "bench"), and the full ncu captures (stage "prof": two ~25 MB reports; gpurun brings back
# at most 64 MiB per call).  Usage: gpurun --timeout 1800 -- 'bash tools/gpu_check.sh tag bench|prof'
TAG=${1:-r01}
STAGE=${2:-bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$STAGE" = bench ]; then
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
ls -la MEASURED_PEAKS.json >> $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py --impl reference --steps 200 --warmup 20 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
cut -c1-300 $OUT/bench_reference.json
timeout 600 python bench.py > $OUT/bench_config3.json 2> $OUT/bench_config3.err; echo "bench rc=$?"
cat $OUT/bench_config3.json
for w in config2 config4 config5 hbm; do
  timeout 600 python bench.py --workload $w --no-cpu --steps 2000 --warmup 600 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  cut -c1-120 $OUT/bench_$w.json; grep -o '"roofline": {[^}]*}' $OUT/bench_$w.json | cut -c1-200; grep -o '"e2e": {"value": [0-9.e+]*' $OUT/bench_$w.json
done
timeout 600 python bench.py --dtype f32 --no-cpu > $OUT/bench_config3_f32.json 2> $OUT/bench_config3_f32.err
cut -c1-120 $OUT/bench_config3_f32.json
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_config3.csv python bench.py --steps 400 --warmup 200 --no-cpu > $OUT/ncu_launches.log 2>&1
else
# full capture of the rollout kernel (config3 and the HBM-resident point)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 \
    -o $OUT/prof_config3 python bench.py --steps 400 --warmup 200 --no-cpu --no-e2e > $OUT/ncu_full_config3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 \
    -o $OUT/prof_hbm python bench.py --workload hbm --steps 40 --warmup 20 --episode-steps 20 --no-cpu --no-e2e > $OUT/ncu_full_hbm.log 2>&1
fi
ls -la $OUT
