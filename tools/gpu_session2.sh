#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_layernorm2d.py tests/test_dwconv.py tests/test_model_parity.py tests/test_gpu_fused.py tests/test_gpu_ss2d.py tests/test_gpu_dct.py -q -m gpu --timeout 600 > gpurun_out/test_gpu_new.log 2>&1; echo "== gpu tests: $(tail -1 gpurun_out/test_gpu_new.log)"
grep -E "^(FAILED|ERROR)|Error|max err" gpurun_out/test_gpu_new.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
TRAMBA_PROJ=conv timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_projconv.json 2> gpurun_out/bench_projconv.err; echo "bench(proj conv) rc=$?"; cut -c1-330 gpurun_out/bench_projconv.json
TRAMBA_LAYERNORM2D=torch timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lntorch.json 2> gpurun_out/bench_lntorch.err; echo "bench(ln torch) rc=$?"; cut -c1-330 gpurun_out/bench_lntorch.json
TRAMBA_DWCONV=torch timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dwtorch.json 2> gpurun_out/bench_dwtorch.err; echo "bench(dw torch) rc=$?"; cut -c1-330 gpurun_out/bench_dwtorch.json
bash tools/gpu_profile.sh
