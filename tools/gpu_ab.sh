#!/bin/bash
# A/B session: parity subset, then bench + warp profile per variant. usage: tools/gpu_ab.sh TAG WORKLOAD "ENV1=.. ENV2=.." "ENV.." ...
TAG=$1; WL=$2; shift 2
mkdir -p gpurun_out
[ -z "$SKIP_TESTS" ] && timeout 900 python -m pytest tests/test_gpu_raytrace.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
i=0
for V in "$@"; do
  echo "== variant $i: [$V]"
  for rep in 1 2; do
    env $V timeout 300 python bench.py --workload $WL --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_v${i}.err | tee gpurun_out/${TAG}_v${i}_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms', round(d['ms_per_step'],4), 'fps', round(d['fps'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['fps'],1), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('samples'))"
  done
  env $V timeout 120 python tools/warp_profile.py $WL > gpurun_out/${TAG}_v${i}_warps.json 2>&1; cut -c1-700 gpurun_out/${TAG}_v${i}_warps.json
  i=$((i+1))
done
