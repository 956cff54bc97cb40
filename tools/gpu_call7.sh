#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_fused test_gpu_fused_sites; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu --timeout 300 > gpurun_out/$f.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/$f.log)"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/$f.log | head -12
done
timeout 300 python tools/fused_ab.py --dtype bf16 > gpurun_out/fused_ab_bf16.jsonl 2>&1; cat gpurun_out/fused_ab_bf16.jsonl
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ss2d_bwd_dx -s 3 -c 1 -o gpurun_out/prof_bwd_dx_enc2 -f python tools/fused_ab.py --dtype bf16 --only enc2 --reps 2 > gpurun_out/ncu_bwd_dx2.log 2>&1; echo "ncu rc=$?"
