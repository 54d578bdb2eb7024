#!/bin/bash
# ad-hoc knob sweep on the GPU box: tools/sweep.sh WORKLOAD  (prints warp profiles for a few settings)
WL=${1:-c2}
for cfg in "fused 16 2"; do
  set -- $cfg
  echo "== path=$1 refill_below=$2 burst=$3"
  B200R_RT_PATH=$1 B200R_REFILL_BELOW=$2 B200R_INNER_BURST=$3 timeout 120 python tools/warp_profile.py $WL 2>&1 | tail -1
  for i in 1 2; do B200R_RT_PATH=$1 B200R_REFILL_BELOW=$2 B200R_INNER_BURST=$3 timeout 200 python bench.py --workload $WL --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', round(d['ms_per_step'],4), 'fps', round(d['fps'],1))"; done
done
