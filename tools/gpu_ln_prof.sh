#!/bin/bash
mkdir -p gpurun_out
prof() {  # name regex site
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -o gpurun_out/prof_$1 -f python tools/side_bench.py --only "$3" --reps 2 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
prof ln_bwd ln2d_bwd_reg "out_norm enc0"
prof ln_bwd2 ln2d_bwd_reg "out_norm enc2"
find gpurun_out -size +20M -print -delete; du -sh gpurun_out
