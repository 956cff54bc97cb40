#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (plain), ncu launch list of the bench command.
mkdir -p gpurun_out
./run_gpu_tests.sh
timeout 600 python -m pytest tests/test_model_parity.py -q -m gpu -x --timeout 300 > gpurun_out/test_model_parity.log 2>&1; echo "== model parity: $(tail -1 gpurun_out/test_model_parity.log)"
grep -E "^(FAILED|ERROR)|Error" gpurun_out/test_model_parity.log | head
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
