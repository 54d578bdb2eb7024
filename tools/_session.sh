mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlaa.py tests/test_gpu_raster.py -m gpu -q -x > gpurun_out/mlaa_pytest.log 2>&1; tail -4 gpurun_out/mlaa_pytest.log
for V in "B200R_X=0" "B200R_MLAA_SCAN=1"; do for WL in c4 c4g; do
env $V timeout 300 python bench.py --workload $WL --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('$V $WL', 'ms', round(d['ms_per_step'],4), 'fps', round(d['fps'],1), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['fps'],1), 'launches', d['gpu_launches'])"
done; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/mlaa2_launches.csv python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=list(csv.reader(open("gpurun_out/mlaa2_launches.csv")))
s=next(i for i,r in enumerate(rows) if r and r[0]=='ID'); h=rows[s]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
d=defaultdict(list)
for r in rows[s+1:]:
    if len(r)>vi: d[(r[ki].split('(')[0][-40:], r[gi])].append(float(r[vi].replace(',','')))
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print(k, len(v), 'avg us', round(sum(v)/len(v)/1e3,1))
PY
