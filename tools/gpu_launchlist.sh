#!/bin/bash
OUT=gpurun_out/${1:-l01}; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $OUT/ncu_launches.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('$OUT/launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
import collections
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki][:60]; v=float(r[vi].replace(',',''))
    agg.setdefault(k,[]).append(v)
for k,v in agg.items(): print(f"{len(v):4d} x {sum(v)/len(v)/1000:9.1f} us  {k}")
PY
