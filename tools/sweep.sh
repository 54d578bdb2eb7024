#!/bin/bash
python -m pytest tests/test_gpu_raytrace.py -m gpu -q -x 2>&1 | tail -2
for B in 2 3 4; do
  echo -n "burst=$B  "
  B200R_INNER_BURST=$B python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['fps'],1))"
done
python tools/warp_profile.py c2
