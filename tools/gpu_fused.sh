#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_fused.py -q -m gpu -x --timeout 600 > gpurun_out/test_gpu_fused.log 2>&1; echo "== fused: $(tail -1 gpurun_out/test_gpu_fused.log)"
grep -E "^(FAILED|ERROR)|Error|error|max err" gpurun_out/test_gpu_fused.log | head -20
timeout -s KILL 600 python -m pytest tests/test_model_parity.py tests/test_gpu_ss2d.py -q -m gpu --timeout 300 > gpurun_out/test_model_parity.log 2>&1; echo "== model parity + ss2d: $(tail -1 gpurun_out/test_model_parity.log)"
grep -E "^(FAILED|ERROR)|Error|max err" gpurun_out/test_model_parity.log | head
timeout -s KILL 300 python tools/kernel_bench.py --dtype bf16 --fused-only > gpurun_out/kb_fused_bf16.log 2>&1; cat gpurun_out/kb_fused_bf16.log | tail -12
timeout -s KILL 300 python tools/kernel_bench.py --dtype f32 --fused-only > gpurun_out/kb_fused_f32.log 2>&1; cat gpurun_out/kb_fused_f32.log | tail -12
