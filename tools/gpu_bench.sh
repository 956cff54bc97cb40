#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --unfused --no-cpu-baseline > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench unfused rc=$?"; cat gpurun_out/bench_unfused.json; tail -3 gpurun_out/bench_unfused.err
