#!/bin/bash
# Round 2, session r03e: deep pipelines again, every slot warmed before the timed region (r03a-d: depth > warm-up frames put the
# first-use cudaMallocs of slots 4.. inside the timed region - that was the "collapse" at depth 6 / 8)
mkdir -p gpurun_out
bash tools/gpu_variants.sh r03e_n1 c2 "B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=4 B200R_E2E_DEPTH=4" "B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=6 B200R_E2E_DEPTH=6" \
  "B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=8 B200R_E2E_DEPTH=8" "B200R_BENCH_DEPTH=6 B200R_E2E_DEPTH=6" "B200R_POOL_CTA_WARPS=4 B200R_BENCH_DEPTH=6 B200R_E2E_DEPTH=5"
S="B200R_BENCH_FAKE_SHARD=8"
bash tools/gpu_variants.sh r03e_s8 c2 "$S B200R_BENCH_DEPTH=6" "$S B200R_BENCH_DEPTH=8" \
  "$S B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=6" "$S B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=8" "$S B200R_POOL_CTA_WARPS=4 B200R_BENCH_DEPTH=8" \
  "$S B200R_POOL_CTA_WARPS=1 B200R_BENCH_DEPTH=8"
S="B200R_BENCH_FAKE_SHARD=2"
bash tools/gpu_variants.sh r03e_s2 c2 "$S B200R_BENCH_DEPTH=4" "$S B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=4" "$S B200R_POOL_CTA_WARPS=2 B200R_BENCH_DEPTH=6"
