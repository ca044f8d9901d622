#!/bin/bash
# full ncu capture of the rollout kernel of the default bench workload
OUT=gpurun_out/${1:-p01}; mkdir -p $OUT
KREGEX=${2:-rollout2_kernel}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 \
    -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra ${@:3} > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log; ls -la $OUT
