#!/bin/bash
# ncu --set full capture of one kernel of a bench workload + launch list.  gpurun -- 'bash tools/gpu_ncu.sh <tag> <workload> <kernel regex>'
TAG=$1; WL=$2; KRN=$3
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$KRN -s 4 -c 1 -f -o gpurun_out/${TAG}_prof_${WL} \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_ncu_${WL}.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/${TAG}_prof_${WL}.ncu-rep
