#!/bin/bash
# GPU iteration: parity tests, then short config3 benches over rollout plans and library variants.
# Usage: bash tools/gpu_sweep.sh tag "G TC;G TC;..." [variant.so ...]
TAG=${1:-q}; PLANS=${2:-"1 16;3 8;5 5"}; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print("%.4g frac=%.4f launch_ms=%.4f clk=%s" % (d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["clocks"]["sm_mhz"]))
'
for lib in default "$@"; do
  [ "$lib" != default ] && export DS_LIB_OVERRIDE=$PWD/$lib || unset DS_LIB_OVERRIDE
  IFS=';' read -ra PL <<< "$PLANS"
  for plan in "${PL[@]}"; do
    set -- $plan
    echo "== lib=$lib plan G=$1 TC=$2 ${WL:-config3}"
    DS_PLAN_G=$1 DS_PLAN_TC=$2 timeout 300 python bench.py --workload ${WL:-config3} --no-cpu --no-e2e --steps 4000 --warmup 600 2>&1 | python -c "$summ"
  done
done
