#!/bin/bash
# Residency experiment on the HBM-resident point (many waves: no quantization): plan x library variant
summ='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print("%.4g frac=%.4f launch_ms=%.4f" % (d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"]))
'
for lib in default "$@"; do
  [ "$lib" != default ] && export DS_LIB_OVERRIDE=$PWD/$lib || unset DS_LIB_OVERRIDE
  for plan in "5 5" "2 8" "1 16"; do
    set -- $plan
    echo "== lib=$lib plan G=$1 TC=$2 hbm"
    DS_PLAN_G=$1 DS_PLAN_TC=$2 timeout 300 python bench.py --workload hbm --no-cpu --no-e2e --steps 80 --warmup 20 --episode-steps 20 2>&1 | python -c "$summ"
  done
done
