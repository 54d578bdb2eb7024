#!/bin/bash
# Short GPU-box session: ray-tracing parity subset, bench without the CPU leg, warp profile.
# usage: tools/gpu_quick.sh TAG WORKLOAD [RT_PATH]
TAG=${1:-q}; WL=${2:-c2}; PTH=${3:-}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_raytrace.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
[ -n "$PTH" ] && export B200R_RT_PATH=$PTH
timeout 300 python bench.py --workload $WL --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${WL}.json"))
print({k:d[k] for k in ['value','fps','ms_per_step']}, 'roofline', round(d['roofline']['frac'],4), 'e2e fps', round(d['e2e']['fps'],1))
PY
timeout 120 python tools/warp_profile.py $WL > gpurun_out/${TAG}_warps_${WL}.json 2>&1; cat gpurun_out/${TAG}_warps_${WL}.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
from collections import defaultdict
rows=list(csv.reader(open("gpurun_out/${TAG}_launches_${WL}.csv")))
s=next(i for i,r in enumerate(rows) if r and r[0]=='ID'); h=rows[s]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
d=defaultdict(list)
for r in rows[s+1:]:
    if len(r)>vi: d[r[ki].split('(')[0][-36:]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print(k, len(v), 'avg us', round(sum(v)/len(v)/1e3,1))
PY
