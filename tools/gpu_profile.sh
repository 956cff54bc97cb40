#!/bin/bash
# ncu evidence for profiles/: (1) launch list of ONE eager step of the bench command (the step bench.py brackets
# with cudaProfilerStart/Stop), (2) --set full of fused backward / forward launches of that step (traffic per launch).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ss2d_bwd -c 33 -o gpurun_out/prof_step_fbwd -f $CMD > gpurun_out/ncu_fbwd_step.log 2>&1
echo "ncu fbwd rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ss2d_fwd -c 33 -o gpurun_out/prof_step_ffwd -f $CMD > gpurun_out/ncu_ffwd_step.log 2>&1
echo "ncu ffwd rc=$?"
ls -la gpurun_out | head -40; du -sh gpurun_out
