#!/bin/bash
# One GPU call of round 2: parity tests first, then the bench lines given as arguments.
#   gpurun --timeout 900 -- 'bash tools/gpu_r02.sh <tag> "<pytest args>" <workload> [<workload> ...]'
TAG=${1:-r02}; shift
PYT=${1:-tests}; shift
mkdir -p gpurun_out
echo "== pytest $PYT"
timeout 600 python -m pytest $PYT -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
for WL in "$@"; do
  echo "== bench $WL"
  timeout 300 python bench.py --workload $WL --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_bench_${WL}.err > gpurun_out/${TAG}_bench_${WL}.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${WL}.json").read())
    print("   value", round(d["value"], 1), d["unit"], "| fps", round(d["fps"], 1), "| serial fps", round(d["serial"]["fps"], 1),
          "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 3), "| e2e fps", round(d["e2e"]["fps"], 1),
          "clk", d["clocks"].get("sm_mhz"))
except Exception as e:
    print("   bench failed:", e); print(open("gpurun_out/${TAG}_bench_${WL}.err").read()[-1500:])
PY
done
