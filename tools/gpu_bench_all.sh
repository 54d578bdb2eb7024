#!/bin/bash
# Bench lines of every BASELINE config on one GPU.  gpurun -- 'bash tools/gpu_bench_all.sh <tag> "<wl>:<steps> ..."'
TAG=$1; shift
mkdir -p gpurun_out
for SPEC in $1; do
  WL=${SPEC%%:*}; ST=${SPEC##*:}
  timeout 600 python bench.py --workload $WL --steps $ST --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_${WL}.err > gpurun_out/${TAG}_bench_${WL}.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${WL}.json").read())
    print("[$WL] value", round(d["value"], 1), d["unit"], "| fps", round(d["fps"], 2), "| serial fps", round(d["serial"]["fps"], 2), "kernel_ms", round(d["roofline"]["kernel_ms"], 4),
          "frac", round(d["roofline"]["frac"], 3), "| e2e fps", round(d["e2e"]["fps"], 2), "| rays/frame", d["config"].get("rays_per_frame"))
except Exception as e:
    print("[$WL] failed:", e); print(open("gpurun_out/${TAG}_bench_${WL}.err").read()[-1500:])
PY
done
