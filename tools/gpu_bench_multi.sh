#!/bin/bash
# multi-GPU arm as the driver launches it: torchrun, one rank per GPU
N=${2:-2}; OUT=gpurun_out/${1:-m01}; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"
tail -3 $OUT/bench_n$N.err
python - <<PY
import json
d=[json.loads(l) for l in open('$OUT/bench_n$N.json') if l.startswith('{')][-1]
print('N',d['n_gpus'],'value',d['value'],'ms/step',d['ms_per_step'],'frac',d['roofline']['frac'])
print('e2e',d['e2e']['value'],d['e2e']['pcie_gbs_this_rank'],'full',d['e2e_full_precision_obs']['value'])
for k,v in d['extra'].items(): print(k,v['n_envs_this_rank'],v['value'],v['roofline']['frac'],v['agg_check'])
print('agg',d['agg_check'])
PY
