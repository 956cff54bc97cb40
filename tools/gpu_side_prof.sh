#!/bin/bash
# side-op microbench + one ncu --set full capture (with source) per kernel of interest
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_layernorm2d.py tests/test_dwconv.py -q -m gpu --timeout 200 2>&1 | tail -3
timeout 200 python tools/side_bench.py --flush read > gpurun_out/side_bench_read.log 2>&1; cat gpurun_out/side_bench_read.log
prof() {  # name regex site
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -o gpurun_out/prof_$1 -f python tools/side_bench.py --only "$3" --reps 2 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
prof ln_fwd ln2d_fwd_reg "out_norm enc0"
prof ln_bwd ln2d_bwd_reg "out_norm enc0"
prof dw3_bwd dwconv_bwd "conv2d enc0"
prof dw3_fwd dwconv_fwd "conv2d enc0"
prof dw7_bwd dwconv_bwd "dwms dec96"
find gpurun_out -size +20M -print -delete; du -sh gpurun_out
