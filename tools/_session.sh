mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/n1.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('N1', {k:round(d[k],4) for k in ('value','fps','ms_per_step')}, round(d['roofline']['kernel_ms'],4), round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['fps'],1), round(d['e2e']['fps_blocking_call'],1), d['clocks'])"; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2> gpurun_out/n2_bench.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('N2', {k:round(d[k],4) for k in ('value','fps','ms_per_step')}, round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['fps'],1)); print(d.get('per_rank'))"
tail -3 gpurun_out/n2_bench.err | cut -c1-300
