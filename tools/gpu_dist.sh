#!/bin/bash
# Multi-GPU session (gpurun --gpus N): parity of the frame pipeline, then bench lines at N GPUs for both assembly modes.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_dist.sh <tag> <N> "<workload:steps> ..." [check-args...]'
TAG=$1; N=$2; SPECS=$3; shift; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/${TAG}_topo.txt
for CHK in "$@"; do
  echo "== dist_check $CHK (N=$N)"
  timeout 400 $TR --master-port 29511 tools/dist_check.py $CHK 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -$((2*N+4)) | tee -a gpurun_out/${TAG}_dist_check_n${N}.log
done
for SPEC in $SPECS; do
  WL=${SPEC%%:*}; ST=${SPEC##*:}
  for MODE in push nccl; do
    echo "== bench $WL N=$N assemble=$MODE"
    B200R_ASSEMBLE=$MODE timeout 600 $TR --master-port 29512 bench.py --gpus $N --workload $WL --steps $ST --warmup 4 --no-cpu-baseline \
        2> gpurun_out/${TAG}_bench_${WL}_n${N}_${MODE}.err > gpurun_out/${TAG}_bench_${WL}_n${N}_${MODE}.json
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${WL}_n${N}_${MODE}.json").read().strip().splitlines()[-1])
    print("   value", round(d["value"], 1), d["unit"], "| fps", round(d["fps"], 1), "| serial fps", round(d["serial"]["fps"], 1),
          "| kernel alone ms", round(d["serial"]["roofline"]["kernel_ms"], 4), "| e2e fps", round(d["e2e"]["fps"], 1),
          "| enqueue ms", round(d["config"]["host_enqueue_ms_per_step"], 4))
except Exception as e:
    print("   bench failed:", e); print(open("gpurun_out/${TAG}_bench_${WL}_n${N}_${MODE}.err").read()[-2000:])
PY
  done
done
