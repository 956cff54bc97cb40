#!/bin/bash
# checkpoint: the driver's GPU test command, smoke, microbenches, configs 4/5, the bench line, the reference arm
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/test_gpu_all.log 2>&1; echo "== gpu tests: $(tail -1 gpurun_out/test_gpu_all.log)"
grep -E "^(FAILED|ERROR)|Error|max err" gpurun_out/test_gpu_all.log | head -20
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
timeout -s KILL 300 python tools/kernel_bench.py --dtype bf16 > gpurun_out/kb_all_bf16.log 2>&1; tail -14 gpurun_out/kb_all_bf16.log | cut -c1-420
timeout -s KILL 200 python tools/side_bench.py --flush read > gpurun_out/side_bench_read.log 2>&1; echo "side_bench rc=$?"
timeout -s KILL 400 python tools/infer_bench.py > gpurun_out/infer_bench.log 2> gpurun_out/infer_bench.err; cat gpurun_out/infer_bench.log; tail -3 gpurun_out/infer_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget-s 60 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
du -sh gpurun_out
