#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_fused test_gpu_fused_sites test_gpu_ss2d test_model_parity; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu --timeout 300 > gpurun_out/$f.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/$f.log)"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/$f.log | head -12
done
timeout 300 python tools/fused_ab.py --dtype bf16 > gpurun_out/fused_ab_bf16.jsonl 2>&1; cat gpurun_out/fused_ab_bf16.jsonl
timeout 300 python tools/fused_ab.py --dtype f32 > gpurun_out/fused_ab_f32.jsonl 2>&1; tail -4 gpurun_out/fused_ab_f32.jsonl
