#!/bin/bash
# final evidence run: bench line (+ CUPTI kernel table), ncu launch list of ONE eager bench step, per-launch DRAM bytes of
# our kernels in that step, one --set full capture (with source) of the fused backward / forward.  gpurun_out/ stays small.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
python tools/launch_summary.py gpurun_out/launches.csv 70 > gpurun_out/launches_summary.txt 2>&1; head -12 gpurun_out/launches_summary.txt
timeout 600 ncu --profile-from-start off -k regex:'ss2d_|ln2d_|dwconv_|dct2d_|merge_skewed|proj_' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/tramba_dram.csv $CMD > gpurun_out/ncu_dram.log 2>&1
echo "ncu dram rc=$?"; wc -l gpurun_out/tramba_dram.csv
python tools/make_traffic_json.py gpurun_out/tramba_dram.csv gpurun_out/roofline_traffic.json | cut -c1-600
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ss2d_bwd -s 3 -c 1 -o gpurun_out/prof_fbwd -f python tools/kernel_bench.py --dtype bf16 --fused-only --only enc0 --reps 2 > gpurun_out/ncu_fbwd.log 2>&1; echo "ncu fbwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ss2d_fwd -s 3 -c 1 -o gpurun_out/prof_ffwd -f python tools/kernel_bench.py --dtype bf16 --fused-only --only enc0 --reps 2 > gpurun_out/ncu_ffwd.log 2>&1; echo "ncu ffwd rc=$?"
find gpurun_out -size +20M -print -delete; du -sh gpurun_out
