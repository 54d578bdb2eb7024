#!/bin/bash
# Round 2, session r03b: CTA size of rt_pool_kernel (8 / 4 / 2 warps), e2e pipeline depth, one rank of 8 emulated.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_raytrace.py -m gpu -x -q -k "scheduling_variants or wavefront or pipelined" 2>&1 | tail -5
bash tools/gpu_variants.sh r03b_n1 c2 "B200R_X=0" "B200R_POOL_CTA_WARPS=4" "B200R_POOL_CTA_WARPS=2" "B200R_E2E_DEPTH=2" "B200R_E2E_DEPTH=4" "B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=3"
bash tools/gpu_variants.sh r03b_s8 c2 "B200R_BENCH_FAKE_SHARD=8" "B200R_BENCH_FAKE_SHARD=8 B200R_POOL_CTA_WARPS=4" "B200R_BENCH_FAKE_SHARD=8 B200R_POOL_CTA_WARPS=2" \
   "B200R_BENCH_FAKE_SHARD=8 B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=6" "B200R_BENCH_FAKE_SHARD=8 B200R_POOL_CTA_WARPS=2 B200R_POOL_TILES_PER_WARP=2" \
   "B200R_BENCH_FAKE_SHARD=8 B200R_POOL_CTA_WARPS=4 B200R_BENCH_DEPTH=6"
for P in 1 8; do timeout 60 python tools/pool_stats.py c2 10 $P; done
bash tools/gpu_variants.sh r03b_c3 c3 "B200R_X=0" "B200R_POOL_CTA_WARPS=4" "B200R_POOL_CTA_WARPS=2"
