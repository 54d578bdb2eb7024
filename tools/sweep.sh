#!/bin/bash
python -m pytest tests/test_gpu_raytrace.py -m gpu -q -x 2>&1 | tail -2
for P in jobs fused generic; do
  echo -n "path=$P  "
  B200R_RT_PATH=$P python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['fps'],1), d['gpu_launches'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/q6_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
from collections import defaultdict
rows=list(csv.reader(open("gpurun_out/q6_launches.csv")))
s=next(i for i,r in enumerate(rows) if r and r[0]=='ID'); h=rows[s]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
d=defaultdict(list)
for r in rows[s+1:]:
    if len(r)>vi: d[r[ki].split('(')[0][-40:]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print(k, len(v), 'avg us', round(sum(v)/len(v)/1e3,1))
PY
