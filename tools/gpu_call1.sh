#!/bin/bash
# Round 2, first GPU call: hardware verdict on everything prepared on the CPU.
mkdir -p gpurun_out
timeout 120 tools/ub/ubench.bin > gpurun_out/ubench.jsonl 2>&1; echo "ubench rc=$?"; cat gpurun_out/ubench.jsonl
bash run_gpu_tests.sh > gpurun_out/run_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -40 gpurun_out/run_gpu_tests.log
timeout 300 python tools/variant_bench.py --dtype bf16 --reps 20 > gpurun_out/variant_bf16.jsonl 2> gpurun_out/variant_bf16.err
echo "variant_bench rc=$?"; tail -3 gpurun_out/variant_bf16.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_fused.py -q -m gpu -x \
    -k "12 or 24" > gpurun_out/sanitizer_${tool}_fused.log 2>&1
  echo "compute-sanitizer $tool fused rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_fused.log | tail -1)"
done
