#!/bin/bash
# Round 2, session r03g: the L2 flush as a lean kernel (2 CTAs per SM) instead of cudaMemsetAsync; 2 tiles per warp for frames in flight
mkdir -p gpurun_out
bash tools/gpu_variants.sh r03g_n1 c2 "B200R_X=0" "B200R_PIPE_FLUSH_MEMSET=1"
S="B200R_BENCH_FAKE_SHARD=8"
bash tools/gpu_variants.sh r03g_s8 c2 "$S B200R_BENCH_DEPTH=6" "$S B200R_BENCH_DEPTH=8" "$S B200R_BENCH_DEPTH=8 B200R_PIPE_FLUSH_MEMSET=1" "$S B200R_BENCH_DEPTH=12" \
  "$S B200R_BENCH_DEPTH=8 B200R_POOL_TILES_PER_WARP=3"
bash tools/gpu_variants.sh r03g_s2 c2 "B200R_BENCH_FAKE_SHARD=2 B200R_BENCH_DEPTH=6" "B200R_BENCH_FAKE_SHARD=4 B200R_BENCH_DEPTH=6"
