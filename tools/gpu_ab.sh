#!/bin/bash
# A/B on ONE box: the build in the tree against libdronestep_base.so (DS_LIB_OVERRIDE), alternating
OUT=gpurun_out/${1:-ab}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_instantiation or golden or returns_recipe or rollout_vs_oracle" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $OUT/pytest_gpu.log
BASE=$PWD/scalable_collision_avoidance_rl_b200/libdronestep_base.so
for w in ${WL:-config3 config2}; do for rep in 1 2 3; do for v in new base; do
E=""; [ $v = base ] && E="DS_LIB_OVERRIDE=$BASE"
env $E timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-e2e --no-extra > $OUT/b_${w}_${v}_$rep.json 2>$OUT/b_${w}_${v}_$rep.err
python -c "import json;d=json.load(open('$OUT/b_${w}_${v}_$rep.json'));print('$w $v $rep',d['roofline']['median_launch_ms'],d['roofline']['min_launch_ms'],round(d['roofline']['frac'],4))"
done; done; done
