#!/bin/bash
# segment-balance scan (DS_RO2_RHO_PERMILLE) per workload on one box
OUT=gpurun_out/${1:-rho}; mkdir -p $OUT
for spec in "config2 55" "config2 90" "config2 130" "config2 180" "config4 55" "config4 25" "config4 0" "config3 55" "config3 70"; do
set -- $spec; w=$1; rho=$2
DS_RO2_RHO_PERMILLE=$rho timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/b_${w}_$rho.json 2>$OUT/b_${w}_$rho.err
python -c "import json;d=json.load(open('$OUT/b_${w}_$rho.json'));print('$w rho$rho',d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],round(d['roofline']['frac'],4))"
done
