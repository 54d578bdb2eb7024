#!/bin/bash
# First GPU call of round 2 (one B200, ~4 min of run time): everything round 1 prepared but could not run.
#   gpurun --timeout 600 -- 'bash tools/gpu_round2_first.sh r02a'
# Every step has its own timeout: the urgent-queue kernel has never run (a hang must not take the box down).
TAG=${1:-r02a}
mkdir -p gpurun_out
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('   value fps', round(d['fps'],1), '| serial fps', round(d['serial']['fps'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), '| e2e fps', round(d['e2e']['fps'],1), 'clk', d['clocks'].get('sm_mhz'))"; }
echo "== 1. full-size AO digests (C3, C5) - first run on a GPU"
B200R_FULLSIZE_GPU=1 timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -4
echo "== 2. urgent-queue kernel: parity first (ray-tracing tests with the variant forced on), then timing"
for V in "B200R_URGENT_T=32" "B200R_URGENT_T=32 B200R_URGENT_NOHIT=1 B200R_URGENT_SHADOW=1"; do
  echo "-- parity [$V]"
  env $V timeout 150 python -m pytest tests/test_gpu_raytrace.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
done
i=0
for V in "B200R_X=0" "B200R_URGENT_T=32" "B200R_URGENT_T=16" "B200R_URGENT_T=32 B200R_URGENT_NOHIT=1" "B200R_URGENT_T=32 B200R_URGENT_SHADOW=1" \
         "B200R_URGENT_T=32 B200R_URGENT_NOHIT=1 B200R_URGENT_SHADOW=1" "B200R_URGENT_T=16 B200R_URGENT_NOHIT=1 B200R_URGENT_SHADOW=1" \
         "B200R_K0_BLOCK=64" "B200R_K0_BLOCK=64 B200R_BENCH_DEPTH=3 B200R_E2E_DEPTH=3"; do
  echo "-- bench [$V]"
  env $V timeout 120 python bench.py --workload c2 --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/${TAG}_v${i}.err | tee gpurun_out/${TAG}_v${i}_bench.json | line || tail -3 gpurun_out/${TAG}_v${i}.err
  i=$((i+1))
done
echo "== 3. warp timeline of the best-looking urgent variant, and two frames in flight on one time axis"
B200R_URGENT_T=32 B200R_URGENT_NOHIT=1 B200R_URGENT_SHADOW=1 timeout 100 python tools/warp_profile.py c2 > gpurun_out/${TAG}_warps_urgent.json 2>&1; cut -c1-600 gpurun_out/${TAG}_warps_urgent.json
timeout 100 python tools/overlap_profile.py c2 2 > gpurun_out/${TAG}_overlap_d2.json 2>&1; cut -c1-900 gpurun_out/${TAG}_overlap_d2.json
echo "== 4. C4: the batched MLAA / resolve walks, first measurement (r01i: 0.635 ms)"
for V in "B200R_X=0" "B200R_MLAA_NOBATCH=1"; do
  env $V timeout 200 python bench.py --workload c4 --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/${TAG}_c4_$(echo $V | tr -c 'A-Za-z0-9' '_').json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   [$V] ms', round(d['ms_per_step'],4), 'fps', round(d['fps'],1))"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_c4.csv \
    python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_profiles.py ${TAG} c4 > /dev/null 2>&1; cat profiles/${TAG}_launches_c4.csv 2>/dev/null | head -14
