#!/bin/bash
# ncu evidence for profiles/ (kept small: gpurun_out/ must stay under 64 MiB or nothing comes back):
# (1) launch list of ONE eager step of the bench command (the step bench.py brackets with cudaProfilerStart/Stop),
# (2) dram bytes of every launch of our kernels in that step, (3) --set full of one fused backward / forward launch.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
python tools/launch_summary.py gpurun_out/launches.csv 70 > gpurun_out/launches_summary.txt 2>&1; head -50 gpurun_out/launches_summary.txt
timeout 600 ncu --profile-from-start off -k regex:tramba --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/tramba_dram.csv $CMD > gpurun_out/ncu_dram.log 2>&1
echo "ncu dram rc=$?"; wc -l gpurun_out/tramba_dram.csv
timeout 300 ncu --set full --clock-control none -k regex:ss2d_bwd -s 3 -c 1 -o gpurun_out/prof_fbwd -f python tools/kernel_bench.py --dtype bf16 --fused-only --only enc0 --reps 2 > gpurun_out/ncu_fbwd.log 2>&1
echo "ncu fbwd rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:ss2d_fwd -s 3 -c 1 -o gpurun_out/prof_ffwd -f python tools/kernel_bench.py --dtype bf16 --fused-only --only enc0 --reps 2 > gpurun_out/ncu_ffwd.log 2>&1
echo "ncu ffwd rc=$?"
find gpurun_out -size +20M -print -delete
du -sh gpurun_out
