#!/bin/bash
# tests (as the driver runs them) + smoke + bench + ncu evidence of the fused path
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/test_gpu_all.log 2>&1; echo "== gpu tests: $(tail -1 gpurun_out/test_gpu_all.log)"
grep -E "^(FAILED|ERROR)|Error|max err" gpurun_out/test_gpu_all.log | head -20
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
timeout -s KILL 300 python tools/kernel_bench.py --dtype bf16 > gpurun_out/kb_all_bf16.log 2>&1; tail -14 gpurun_out/kb_all_bf16.log
TRAMBA_LINEAR2D=conv timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conv.json 2> gpurun_out/bench_conv.err; echo "bench(conv) rc=$?"; cut -c1-400 gpurun_out/bench_conv.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
bash tools/gpu_profile.sh
