#!/bin/bash
# The C++ host program on N GPUs: frames must equal the single-GPU run (last frame dumped by both), both assembly modes.
N=$1; M=oracle/_ref/models/chessboard.tri
cd /root/repo
./renderer_b200/b200renderer -b -n 12 -m 9 --width 1920 --height 1080 --no-reflections --dump gpurun_out/cli_one --frames 11 $M | tail -1
for A in push nccl; do
  ./renderer_b200/b200renderer -b -n 12 -m 9 --width 1920 --height 1080 --no-reflections --gpus $N --assemble $A --dump gpurun_out/cli_$A $M | tail -1
  cmp gpurun_out/cli_one_11.xrgb gpurun_out/cli_${A}_11.xrgb && echo "   --gpus $N --assemble $A: last frame identical to the single-GPU run"
done
./renderer_b200/b200renderer -b -n 300 -m 9 --width 1920 --height 1080 --no-reflections --frames-in-flight 4 --gpus $N $M | tail -1
./renderer_b200/b200renderer -b -n 300 -m 9 --width 1920 --height 1080 --no-reflections $M | tail -1
rm -f gpurun_out/cli_*.xrgb
