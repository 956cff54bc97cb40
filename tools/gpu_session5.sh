#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_proj_gather.py tests/test_layernorm2d.py tests/test_dwconv.py tests/test_model_parity.py tests/test_gpu_fused.py tests/test_gpu_ss2d.py -q -m gpu --timeout 600 > gpurun_out/test_gpu_new.log 2>&1; echo "== gpu tests: $(tail -1 gpurun_out/test_gpu_new.log)"
grep -E "^(FAILED|ERROR)|Error|max err" gpurun_out/test_gpu_new.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
TRAMBA_PROJ_GATHER=torch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pgtorch.json 2> gpurun_out/bench_pgtorch.err; echo "bench(pg torch) rc=$?"; cut -c1-330 gpurun_out/bench_pgtorch.json
du -sh gpurun_out
