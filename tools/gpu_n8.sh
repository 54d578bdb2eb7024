#!/bin/bash
# The 8-GPU session of round 2: pipeline parity at N=8 (C2, C2+MLAA, C5), bench lines C2 (both assembly modes) and C5, the CLI.
bash tools/gpu_dist.sh r02q 8 "c2:100 c5:16" c2 "c2 mlaa" c5
for D in 2 8; do
  echo "== bench c2 N=8 push, $D frames in flight per rank"
  B200R_BENCH_DEPTH=$D timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload c2 --steps 100 --warmup 4 --no-cpu-baseline 2>/dev/null > gpurun_out/r02q_bench_c2_n8_push_d$D.json
  python -c "
import json; d=json.loads(open('gpurun_out/r02q_bench_c2_n8_push_d$D.json').read().strip().splitlines()[-1]); print('   fps', round(d['fps'],1), 'serial', round(d['serial']['fps'],1), 'e2e', round(d['e2e']['fps'],1))"
done
echo "== bench c2 N=8 push, no L2 flush (experiment)"
B200R_BENCH_FLUSH=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload c2 --steps 100 --warmup 4 --no-cpu-baseline 2>/dev/null > gpurun_out/r02q_bench_c2_n8_push_noflush.json
python -c "
import json; d=json.loads(open('gpurun_out/r02q_bench_c2_n8_push_noflush.json').read().strip().splitlines()[-1]); print('   fps', round(d['fps'],1), 'serial', round(d['serial']['fps'],1), 'e2e', round(d['e2e']['fps'],1))"
bash tools/gpu_cli_dist.sh 8 2>&1 | tee gpurun_out/r02q_cli_n8.log
