#!/bin/bash
WL=${1:-c2}
for m in 4 8 16 24; do
  echo "== donate min idle=$m"
  B200R_DONATE_MIN_IDLE=$m timeout 120 python tools/warp_profile.py $WL 2>&1 | tail -1 | cut -c1-500
  for i in 1 2; do B200R_DONATE_MIN_IDLE=$m timeout 200 python bench.py --workload $WL --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],4), 'fps', round(d['fps'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],4))"; done
done
