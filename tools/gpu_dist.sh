#!/bin/bash
# N-GPU session (gpurun --gpus N): pipelined frame assembly parity check, then bench --gpus N per frames-in-flight setting.
# usage: tools/gpu_dist.sh TAG N DEPTH...
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check.log 2>&1; grep dist_check gpurun_out/${TAG}_dist_check.log || tail -20 gpurun_out/${TAG}_dist_check.log
port=29520
for D in "$@"; do
  port=$((port+1))
  B200R_BENCH_DEPTH=$D B200R_E2E_DEPTH=$D timeout 300 $TR --master-port $port bench.py --gpus $N --steps 100 --warmup 5 2> gpurun_out/${TAG}_n${N}_d${D}.err | grep '^{' | tee gpurun_out/${TAG}_n${N}_d${D}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=$N depth $D: value fps', round(d['fps'],1), 'ms', round(d['ms_per_step'],4), '| serial fps', round(d['serial']['fps'],1), 'frac', round(d['roofline']['frac'],3), '| e2e fps', round(d['e2e']['fps'],1), '| launches', d['gpu_launches'], 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
print('   per rank', {k:[round(x,4) for x in v] for k,v in d.get('per_rank',{}).items()})" || tail -15 gpurun_out/${TAG}_n${N}_d${D}.err
done
