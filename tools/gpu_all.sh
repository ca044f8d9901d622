#!/bin/bash
# All BASELINE configs, device-resident timing only (no e2e / cpu legs). Usage: bash tools/gpu_all.sh tag [extra bench args]
TAG=${1:-all}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for w in config2 config3 config4 config5; do
  timeout 600 python bench.py --workload $w --no-cpu --no-e2e --steps 2000 --warmup 600 "$@" > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python - "$OUT/bench_$w.json" "$w" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%s %s %.4g frac=%.4f launch_ms=%.4f" % (sys.argv[2], d["dtype"], d["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"]))
except Exception as e: print(sys.argv[2], "failed", e)
PY
done
