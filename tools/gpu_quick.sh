#!/bin/bash
# Quick GPU iteration: parity tests + short benches over rollout plans.  Usage: bash tools/gpu_quick.sh tag
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
for plan in "3 8" "5 5" "6 4" "12 2" "25 1" "1 16" "2 12"; do
  set -- $plan
  echo "== plan G=$1 TC=$2"
  DS_PLAN_TCMAX=16 DS_PLAN_G=$1 DS_PLAN_TC=$2 timeout 300 python bench.py --no-cpu --no-e2e --steps 4000 --warmup 600 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks'])
"
done
for w in config2 config4 config5; do
  echo "== $w"; timeout 600 python bench.py --workload $w --no-cpu --no-e2e --steps 2000 --warmup 600 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; cat $OUT/bench_$w.json | cut -c1-200; tail -2 $OUT/bench_$w.err
done
echo "== f32"; timeout 300 python bench.py --dtype f32 --no-cpu --no-e2e --steps 4000 > $OUT/bench_f32.json 2>&1; cut -c1-200 $OUT/bench_f32.json
echo "== logmode1"; timeout 300 python bench.py --log-mode 1 --no-cpu --no-e2e --steps 4000 > $OUT/bench_lm1.json 2>&1; cut -c1-200 $OUT/bench_lm1.json
