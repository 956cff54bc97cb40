#!/bin/bash
# The ncu launch list of ONE eager training step of bench.py (the recipe's `--metrics gpu__time_duration.sum
# --clock-control none` pass: serialised, cold-cache per-launch times -- compare SHARES with the CUPTI table), and the
# per-launch DRAM bytes of our kernels (-> profiles/roofline_traffic.json = bench.py's roofline.traffic).
#   gpurun --timeout 1500 -- 'bash tools/gpu_launchlist.sh'
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph \
  > gpurun_out/launchlist_bench.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_step.csv > gpurun_out/r02_launches_step_summary.txt 2>&1; head -30 gpurun_out/r02_launches_step_summary.txt
timeout 900 ncu --profile-from-start off -k regex:'ss2d_|ln2d_|dwconv_|dct2d_|batch_sum|proj_' \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r02_tramba_dram_per_launch.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph \
  > gpurun_out/traffic_bench.log 2>&1; echo "traffic rc=$?"
python tools/make_traffic_json.py gpurun_out/r02_tramba_dram_per_launch.csv gpurun_out/roofline_traffic.json
rm -f gpurun_out/r02_launches_step.csv   # several MB; the summary is what is kept
