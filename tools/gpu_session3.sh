#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_layernorm2d.py tests/test_dwconv.py tests/test_model_parity.py tests/test_gpu_fused.py tests/test_gpu_ss2d.py -q -m gpu --timeout 600 > gpurun_out/test_gpu_new.log 2>&1; echo "== gpu tests: $(tail -1 gpurun_out/test_gpu_new.log)"
grep -E "^(FAILED|ERROR)|Error|max err" gpurun_out/test_gpu_new.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout -s KILL 200 python tools/kernel_bench.py --dtype bf16 --fused-only > gpurun_out/kb_fused_bf16.log 2>&1; tail -10 gpurun_out/kb_fused_bf16.log
du -sh gpurun_out
