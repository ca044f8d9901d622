#!/bin/bash
# Round-2 record: everything profiles/r02 cites.  Usage (from the build container):
#   tools/dev_cycle.sh r02 'bash tools/gpu_r02_record.sh r02 bench' 2400
#   tools/dev_cycle.sh r02p 'bash tools/gpu_r02_record.sh r02p prof' 1800
TAG=${1:-r02}; STAGE=${2:-bench}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$STAGE" = bench ]; then
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_config3.json 2> $OUT/bench_config3.err; echo "bench rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --log-mode 2 --no-cpu --no-extra > $OUT/bench_config3_logrcp.json 2> $OUT/bench_config3_logrcp.err
DS_RO2=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-extra --no-e2e > $OUT/bench_config3_round1_kernel.json 2> $OUT/bench_config3_round1_kernel.err
for w in config2 config4 config5; do
  timeout 600 python bench.py --workload $w --no-cpu --no-extra --steps 20 --warmup 5 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
timeout 600 python bench.py --workload hbm --no-cpu --no-extra --no-e2e --steps 6 --warmup 3 > $OUT/bench_hbm.json 2> $OUT/bench_hbm.err
timeout 600 python bench.py --dtype f32 --no-cpu --no-extra --steps 20 --warmup 5 > $OUT/bench_config3_f32.json 2> $OUT/bench_config3_f32.err
for f in config3 config3_logrcp config3_round1_kernel config2 config4 config5 hbm config3_f32; do
python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_$f.json'))
    e=d.get('e2e') or {}
    print('$f', 'value %.4g'%d['value'], 'launch ms %.4f'%d['roofline']['median_launch_ms'], 'frac %.4f'%d['roofline']['frac'], 'e2e %.4g'%(e.get('value') or 0))
except Exception as ex: print('$f', 'FAILED', ex)
PY
done
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_config3.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $OUT/ncu_launches.log 2>&1
elif [ "$STAGE" = prof ]; then
# one full capture per call (gpurun brings back at most 64 MiB): $3 = config3 | config4 | hbm
W=${3:-config3}
EXTRA=""; [ "$W" = hbm ] && EXTRA="--steps 2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout2_kernel -s 3 -c 1 \
    -o $OUT/prof_$W python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra $EXTRA > $OUT/ncu_full_$W.log 2>&1
tail -2 $OUT/ncu_full_$W.log
else
# memcheck + racecheck + synccheck over the rollout tests (small cases)
for tool in memcheck racecheck synccheck; do
timeout 1200 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rollout_vs_oracle_and_stepping or dense_and_large or returns_recipe or fused_with_observation" > $OUT/sanitizer_$tool.log 2>&1
echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_$tool.log | tail -3
done
fi
ls $OUT | wc -l
