#!/bin/bash
# Round-end GPU session on one B200: all parity tests, smoke, bench (C2 with the CPU arm, C4), warp timeline, ncu launch lists,
# one ncu --set full capture of the dominant kernel.  usage (under gpurun): bash tools/gpu_final.sh TAG
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc > gpurun_out/${TAG}_nproc.txt
timeout 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py --workload c2 --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; cut -c1-1500 gpurun_out/${TAG}_bench_c2.json
timeout 400 python bench.py --workload c4 --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; cut -c1-600 gpurun_out/${TAG}_bench_c4.json
timeout 120 python tools/warp_profile.py c2 > gpurun_out/${TAG}_warps_c2.json 2>&1; cut -c1-400 gpurun_out/${TAG}_warps_c2.json
for WL in c2 c4; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches_${WL}.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rt_primary_kernel -s 4 -c 2 -f -o gpurun_out/${TAG}_prof_c2 \
    python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
