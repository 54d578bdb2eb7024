#!/bin/bash
# Pipelined-protocol A/B on one GPU: ray-tracing parity tests, then bench per env variant.
# usage: tools/gpu_pipe.sh TAG WORKLOAD "ENV.." "ENV.." ...
TAG=$1; WL=$2; shift 2
mkdir -p gpurun_out
[ -z "$SKIP_TESTS" ] && { timeout 600 python -m pytest tests/test_gpu_raytrace.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log; }
i=0
for V in "$@"; do
  echo "== variant $i: [$V]"
  env $V timeout 300 python bench.py --workload $WL --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_v${i}.err | tee gpurun_out/${TAG}_v${i}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('   value fps', round(d['fps'],1), 'ms', round(d['ms_per_step'],4), 'in flight', d['config'].get('frames_in_flight'), 'enqueue_ms', round(d['config'].get('host_enqueue_ms_per_step') or 0,4), '| serial fps', round(d['serial']['fps'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), '| e2e fps', round(d['e2e']['fps'],1), 'blocking', round(d['e2e'].get('fps_blocking_call',0),1), '| launches', d['gpu_launches'], 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))"
  tail -2 gpurun_out/${TAG}_v${i}.err
  i=$((i+1))
done
