#!/bin/bash
# C4 (rasteriser) A/B: raster + MLAA parity tests, bench per env variant, ncu launch list of the first variant.
# usage: tools/gpu_c4ab.sh TAG WORKLOAD "ENV.." "ENV.." ...
TAG=$1; WL=$2; shift 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_raster.py tests/test_gpu_mlaa.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
i=0
for V in "$@"; do
  echo "== variant $i: [$V]"
  env $V timeout 300 python bench.py --workload $WL --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_v${i}.err | tee gpurun_out/${TAG}_v${i}_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms', round(d['ms_per_step'],4), 'fps', round(d['fps'],1), 'e2e', round(d['e2e']['fps'],1), 'blocking', round(d['e2e'].get('fps_blocking_call',0),1), 'clk', d['clocks'].get('sm_mhz'))"
  env $V timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_v${i}_launches.csv \
      python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
from collections import defaultdict
rows=list(csv.reader(open("gpurun_out/${TAG}_v${i}_launches.csv")))
s=next(i for i,r in enumerate(rows) if r and r[0]=='ID'); h=rows[s]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
d=defaultdict(list)
for r in rows[s+1:]:
    if len(r)>vi: d[r[ki].split('(')[0][-40:]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print('   ', k, len(v), 'avg us', round(sum(v)/len(v)/1e3,1))
PY
  i=$((i+1))
done
