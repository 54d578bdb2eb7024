#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list + full capture of the dominant kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_session.sh <tag> [workload]
TAG=${1:-r01}; WL=${2:-c2}; KRN=${3:-rt_primary_kernel}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc > gpurun_out/${TAG}_nproc.txt; lscpu | head -20 >> gpurun_out/${TAG}_nproc.txt
[ -z "$SKIP_TESTS" ] && { timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --workload $WL --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err; cat gpurun_out/${TAG}_bench_${WL}.json
[ "$KRN" = rt_primary_kernel ] && timeout 200 python tools/tile_profile.py $WL > gpurun_out/${TAG}_tiles_${WL}.json 2>&1; cat gpurun_out/${TAG}_tiles_${WL}.json
[ "$KRN" = rt_primary_kernel ] && timeout 200 python tools/warp_profile.py $WL > gpurun_out/${TAG}_warps_${WL}.json 2>&1; cat gpurun_out/${TAG}_warps_${WL}.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$KRN -s 4 -c 3 -f -o gpurun_out/${TAG}_prof_${WL} \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
