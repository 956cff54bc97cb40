#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/kernel_bench.py --dtype bf16 --fused-only --only enc0 --reps 3"
$CMD > gpurun_out/plain_kb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ss2d_fwd -s 3 -c 1 -o gpurun_out/prof_ffwd -f $CMD > gpurun_out/ncu_ffwd.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ffwd.log
ncu --set full --clock-control none --import-source on -k regex:ss2d_bwd -s 3 -c 1 -o gpurun_out/prof_fbwd -f $CMD > gpurun_out/ncu_fbwd.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_fbwd.log
