#!/bin/bash
# dev loop (build container): rebuild libdronestep.so, keep a copy of it next to the results (so that
# ncu reports can be joined with the right SASS later), then run a GPU script through gpurun.
# usage: tools/dev_cycle.sh TAG "gpu command using \$TAG" [gpurun timeout]
set -e -o pipefail
cd /root/repo
TAG=$1; CMD=$2; TMO=${3:-1500}
python -c "import __graft_entry__ as g; g.build()" | tail -1
mkdir -p gpurun_out/$TAG
cp scalable_collision_avoidance_rl_b200/libdronestep.so gpurun_out/$TAG/lib.so
ls -la scalable_collision_avoidance_rl_b200/libdronestep.so | awk '{print $6, $7, $8}'
/usr/local/graft/bin/gpurun --timeout $TMO -- "$CMD" 2>&1 | tail -${4:-20}
