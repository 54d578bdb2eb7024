#!/bin/bash
# CTA-size x frames-in-flight sweep of the pipelined protocol (developer builds from tools/build_variant.sh).
TAG=$1; shift
mkdir -p gpurun_out
B200R_LIB=$PWD/renderer_b200/libb200render_b32.so timeout 300 python -m pytest tests/test_gpu_raytrace.py -m gpu -q -x > gpurun_out/${TAG}_pytest_b32.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_b32.log
for L in "" _b128 _b64 _b32; do
  for D in "$@"; do
    B200R_LIB=$PWD/renderer_b200/libb200render$L.so B200R_BENCH_DEPTH=$D B200R_E2E_DEPTH=$D timeout 300 python bench.py --workload c2 --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}$L_d$D.err | tee gpurun_out/${TAG}${L}_d${D}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('lib[$L] depth $D: value fps', round(d['fps'],1), '| serial fps', round(d['serial']['fps'],1), 'frac', round(d['roofline']['frac'],3), '| e2e fps', round(d['e2e']['fps'],1), 'blocking', round(d['e2e'].get('fps_blocking_call',0),1), 'clk', d['clocks'].get('sm_mhz'))"
  done
done
