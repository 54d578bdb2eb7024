#!/bin/bash
# A/B of environment variants on one bench workload.  gpurun -- 'bash tools/gpu_variants.sh <tag> <workload> "VAR=1 VAR2=x" "..." ...'
TAG=$1; WL=$2; shift; shift
mkdir -p gpurun_out
i=0
for V in "$@"; do
  env $V timeout 200 python bench.py --workload $WL --steps 40 --warmup 4 --no-cpu-baseline 2> gpurun_out/${TAG}_v${i}.err > gpurun_out/${TAG}_v${i}.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_v${i}.json").read())
    print("[$V] fps", round(d["fps"], 1), "| serial fps", round(d["serial"]["fps"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 4),
          "frac", round(d["roofline"]["frac"], 3), "| e2e fps", round(d["e2e"]["fps"], 1))
except Exception as e:
    print("[$V] failed:", e); print(open("gpurun_out/${TAG}_v${i}.err").read()[-800:])
PY
  i=$((i+1))
done
