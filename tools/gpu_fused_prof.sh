#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ss2d_bwd -s 3 -c 1 -o gpurun_out/prof_fbwd -f python tools/kernel_bench.py --dtype bf16 --fused-only --only enc0 --reps 2 > gpurun_out/ncu_fbwd.log 2>&1; echo "ncu fbwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ss2d_fwd -s 3 -c 1 -o gpurun_out/prof_ffwd -f python tools/kernel_bench.py --dtype bf16 --fused-only --only enc0 --reps 2 > gpurun_out/ncu_ffwd.log 2>&1; echo "ncu ffwd rc=$?"
find gpurun_out -size +20M -print -delete; du -sh gpurun_out
